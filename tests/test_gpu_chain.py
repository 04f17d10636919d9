"""GPU parity: the full receive chain of TB/Simulated/testFullDemodChain.cs through the C ABI vs the
oracle blocks wired exactly as that test wires them (per-sample FLL -> matched filter, Mueller-Muller and Costas once
per 4096-sample frame, three ZMQ topics per frame)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FS = 10_000_000
RS = FS // 30
FRAME = 4096          # samplesPerFrame :18


def _bits(n, seed):
    rng = np.random.default_rng(seed)
    return "".join("01"[b] for b in rng.integers(0, 2, n))


def _impaired_signal(orc, n_bits=4096, seed=5, repeats=1):
    """modulator.Modulate(4096 random bits) + (-90 dBFS noise), times the two unstable LOs (:44-73)."""
    x = orc.QPSKModulator(FS, RS, 0.9, 10).Modulate(_bits(n_bits, seed))
    x = np.tile(x, repeats)
    tx = orc.NCO(935e6, FS, 20, 120, seed=seed, stream=0)
    rx = orc.NCO(935e6, FS, 10, 30, seed=seed, stream=1)
    noise = orc.noise_iq(-90.0, x.size // 2, seed, 2)
    return orc.channel_apply(tx, rx, 1, x, noise)


def _oracle_chain(orc):
    sps = FS // RS
    bn, zeta = 0.000000002, 1.0 / np.sqrt(2.0)
    om = 2.0 * np.pi * bn
    rrc = orc.ComplexFIRFilter(orc.real_taps_to_iq(orc.RRCFilter.generateCoefficents(11, .9, FS, RS)))
    fll = orc.FLLBandEdgeFilter(float(sps), float(np.float32(.9)), 10, float(np.float32(0.1)))
    mm = orc.MuellerMuller(float(sps), 2.0 * zeta * om / 1.0, om * om / 1.0)
    costas = orc.CostasLoopQpsk(float(RS), float(RS // 10))
    return rrc, fll, mm, costas


def _oracle_frames(orc, y):
    """The loop body of testFullDemodChain.cs:62-112 over the signal y, frame by frame -> list of (topic, bytes)."""
    rrc, fll, mm, costas = _oracle_chain(orc)
    out = []
    for a in range(0, y.size, 2 * FRAME):
        blk = y[a:a + 2 * FRAME]
        pre = fll.Process(blk)                      # per-sample FLL (:73) == block call (state carried)
        mf = rrc.Filter(pre)                        # per-sample Filter (:79) == block call
        decided = mm.Process(mf)                    # :84
        out.append((b"baseband", pre.astype("<f4").tobytes()))
        if decided.size:
            out.append((b"baseband_PostSymbolSync", decided.astype("<f4").tobytes()))
            out.append((b"baseband_PostSymbolSyncPostCostas", costas.Process(decided).astype("<f4").tobytes()))
    return out


def test_default_params_are_the_tests_literals(gpu):
    ch = gpu.FullDemodChain()
    p = ch.params
    assert (p.sample_rate, p.symbol_rate) == (FS, RS)
    assert (p.rrc_span, p.rrc_alpha) == (11.0, 0.9)
    assert (p.fll_sps, p.fll_size) == (30.0, 10)
    assert p.fll_rolloff == np.float32(0.9) and p.fll_bw == np.float32(0.1)
    om = 2.0 * np.pi * 0.000000002
    assert p.mm_sps == 30.0 and p.mm_kp == 2.0 * (1.0 / np.sqrt(2.0)) * om and p.mm_ki == om * om
    assert (p.costas_sample_rate, p.costas_bw_hz, p.costas_damping) == (float(RS), float(RS // 10), 0.707)


def test_zmq_frames_bit_identical_to_oracle_wiring(gpu, orc):
    y = _impaired_signal(orc)
    want = _oracle_frames(orc, y)
    ch = gpu.FullDemodChain()
    ch.set_fir_mode(gpu.FIR_EXACT)
    got = ch.zmq_frames(y, FRAME)
    assert [t for t, _ in got] == [t for t, _ in want]
    for (t, g), (_, w) in zip(got, want):
        assert g == w, t
    # the symbol topics carry about samples/sps symbols in total
    nsym = sum(len(b) for t, b in got if t == b"baseband_PostSymbolSync") // 8
    assert abs(nsym - (y.size // 2) / 30) <= 3


def test_one_shot_equals_framed_and_fast_mode_within_tolerance(gpu, orc):
    y = _impaired_signal(orc, seed=9)
    rrc, fll, mm, costas = _oracle_chain(orc)
    pre = fll.Process(y)
    sym = mm.Process(rrc.Filter(pre))
    cos = costas.Process(sym)
    ex = gpu.FullDemodChain()
    ex.set_fir_mode(gpu.FIR_EXACT)
    bb, sy, co = ex.Process(y)
    assert np.array_equal(bb.view(np.uint32), pre.view(np.uint32))
    assert np.array_equal(sy.view(np.uint32), sym.view(np.uint32))
    assert np.array_equal(co.view(np.uint32), cos.view(np.uint32))
    # default (FAST matched filter): the stated fp32 tolerance
    fa = gpu.FullDemodChain()
    bb2, sy2, co2 = fa.Process(y)
    assert np.array_equal(bb2.view(np.uint32), pre.view(np.uint32))      # the FLL does not depend on the MF mode
    assert sy2.shape == sym.shape
    assert np.abs(sy2 - sym).max() <= 1e-5 * np.abs(sym).max()
    assert np.abs(co2 - cos).max() <= 1e-5 * np.abs(cos).max()
    st_o = (fll.state, costas.GetState())
    st = ex.loop_state()
    assert st["fll_phase"] == st_o[0][0] and st["fll_freq"] == st_o[0][1]
    # fp64 loop state: the device sin/cos are within 1 ulp of glibc's (DESIGN.md "Transcendentals"), outputs are fp32
    assert abs(st["costas_theta"] - st_o[1][0]) <= 1e-12 * max(1.0, abs(st_o[1][0]))
    assert abs(st["costas_freq"] - st_o[1][1]) <= 1e-12 * max(1.0, abs(st_o[1][1]))


def test_batch_channels_and_device_entry(gpu, orc):
    import torch
    C = 5
    ys = [_impaired_signal(orc, n_bits=1024, seed=20 + c) for c in range(C)]
    n = ys[0].size
    ch = gpu.FullDemodChain(channels=C)
    ch.set_fir_mode(gpu.FIR_EXACT)
    x = torch.from_numpy(np.stack(ys)).cuda()
    cap = ch.symbols_bound(n)
    bb = torch.zeros((C, n), dtype=torch.float32, device="cuda")
    sy = torch.zeros((C, cap), dtype=torch.float32, device="cuda")
    co = torch.zeros((C, cap), dtype=torch.float32, device="cuda")
    ns = torch.zeros(C, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    half = (n // 4) * 2
    # two calls (state carried), second one on a strided view of the same buffers
    ch.process_dev(x.data_ptr(), half, n, bb.data_ptr(), n, sy.data_ptr(), cap, co.data_ptr(), cap, ns.data_ptr())
    torch.cuda.synchronize()
    n1 = ns.cpu().numpy().copy()
    sy1, co1 = sy.cpu().numpy().copy(), co.cpu().numpy().copy()
    ch.process_dev(x.data_ptr() + 4 * half, n - half, n, bb.data_ptr() + 4 * half, n, sy.data_ptr(), cap, co.data_ptr(), cap,
                   ns.data_ptr())
    torch.cuda.synchronize()
    n2 = ns.cpu().numpy()
    for c in range(C):
        # the oracle is fed the same two chunks: Mueller-Muller's mu = newTime - floor(newTime) (MuellerMuller.cs:113-115)
        # rounds differently when baseIndex restarts from a freshly trimmed buffer, so chunking is visible at 1 ulp
        rrc, fll, mm, costas = _oracle_chain(orc)
        pre = fll.Process(ys[c])
        mf = rrc.Filter(pre)
        sym = np.concatenate([mm.Process(mf[:half]), mm.Process(mf[half:])])
        cos = costas.Process(sym)
        assert np.array_equal(bb[c].cpu().numpy().view(np.uint32), pre.view(np.uint32))
        got_sym = np.concatenate([sy1[c, : 2 * n1[c]], sy[c].cpu().numpy()[: 2 * n2[c]]])
        got_cos = np.concatenate([co1[c, : 2 * n1[c]], co[c].cpu().numpy()[: 2 * n2[c]]])
        assert np.array_equal(got_sym.view(np.uint32), sym.view(np.uint32))
        assert np.array_equal(got_cos.view(np.uint32), cos.view(np.uint32))


def test_chain_error_behaviour(gpu):
    ch = gpu.FullDemodChain()
    with pytest.raises(gpu.ArgumentException):
        ch.Process(np.zeros(5, np.float32))                       # odd interleaved length
    bb, sy, co = ch.Process(np.zeros(0, np.float32))
    assert bb.size == 0 and sy.size == 0 and co.size == 0
    with pytest.raises(gpu.ArgumentOutOfRangeException):
        gpu.FullDemodChain(fll_size=0)                            # Band-Edge Filter.cs:44
    with pytest.raises(gpu.QpskCudaError):
        gpu.FullDemodChain(mm_sps=0.0)                            # unsupported: the reference loop would never advance
    with pytest.raises(TypeError):
        gpu.FullDemodChain(not_a_param=1)
