"""GPU parity of the demodulator with the matched filter fused into the symbol-stage kernel (symsync_decode_kernel<DIFF, MFW>,
the default path: QPSK_FIR_EXACT, real taps, <= 65 taps).  Everything is compared with the oracle bit for bit: bits,
constellation points (which go through the SEPARATE matched-filter kernel on the same delay line) and loop state."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TSC = "11001010011101100100100110101100" + "01110100111001011010001101101001"


def _stream(orc, fs, rs, alpha, span, n_bursts, n_payload, seed, tsc=None, noise=-45.0):
    rng = np.random.default_rng(seed)
    mod = orc.QPSKModulator(fs, rs, alpha, span, True, tsc)
    tx, rx = orc.NCO(100e6, fs, 1, seed=seed, stream=0), orc.NCO(100e6, fs, 1, seed=seed, stream=1)
    parts = []
    for b in range(n_bursts):
        s = mod.ModulateBytes(rng.integers(0, 256, n_payload, dtype=np.uint8).tobytes(), b"S", b"E")
        nz = orc.noise_iq(noise, s.size // 2, seed, 2, first_sample=b * 100000)
        parts.append(orc.channel_apply(tx, rx, 1, s, nz))
    return np.concatenate(parts)


@pytest.mark.parametrize("fs,rs,span,alpha", [(10_000_000, 5_000_000, 10, 0.4), (4000, 1000, 10, 0.35), (8000, 1000, 8, 0.5),
                                               (3000, 1000, 6, 0.9), (16000, 1000, 6, 0.35)])
def test_fused_mf_bits_any_chunking(gpu, orc, fs, rs, span, alpha):
    """21 / 41 / 65 / 19 taps run fused (2 or 4 matched-filter warps), 97 taps (sps 16) take the separate exact kernel;
    chunk cuts at odd positions, shorter than a round (32 samples) and shorter than the filter."""
    alpha = float(np.float32(alpha))
    x = _stream(orc, fs, rs, alpha, span, 3, 120, seed=fs % 97)
    L = x.size // 2
    kw = dict(RrcAlpha=alpha, rrcSpan=span, SymbolSyncBandwith=0.002, tsc=None)
    rng = np.random.default_rng(3)
    for trial in range(3):
        cuts = [0, L] if trial == 0 else sorted(set([0, L, 5, 31, 33, 64, 1000] + [int(v) for v in rng.integers(0, L, 12)]))
        od, gd = orc.QPSKDeModulator(fs, rs, **kw), gpu.QPSKDeModulator(fs, rs, **kw)
        for a, b in zip(cuts[:-1], cuts[1:]):
            assert gd.DeModulate(x[2 * a:2 * b]) == od.DeModulate(x[2 * a:2 * b]), (trial, a, b)
        ws, gs = od.loop_state(), gd.loop_state()
        for k in ("costas_theta", "costas_freq", "mm_mu", "mm_integral"):
            assert abs(ws[k] - gs[k]) <= 1e-5 * max(1.0, abs(ws[k])), k


def test_fused_mf_shares_the_delay_line_with_the_separate_filter(gpu, orc):
    """DeModulate (fused matched filter) and deModulateConstellation (separate matched-filter kernel) alternate on one
    handle: both advance the same delay line, MM queue and Costas state, exactly as the reference object's calls do."""
    fs, rs = 4000, 1000
    x = _stream(orc, fs, rs, 0.35, 10, 4, 90, seed=8)
    kw = dict(RrcAlpha=0.35, rrcSpan=10, SymbolSyncBandwith=0.002, tsc=None)
    od, gd = orc.QPSKDeModulator(fs, rs, **kw), gpu.QPSKDeModulator(fs, rs, **kw)
    L = x.size // 2
    cuts = [0, 700, 701, 1500, 2222, 2240, 3001, L]
    for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        seg = x[2 * a:2 * b]
        if i % 2 == 0:
            assert gd.DeModulate(seg) == od.DeModulate(seg), i
        else:
            w, g = od.deModulateConstellation(seg), gd.deModulateConstellation(seg)
            assert np.array_equal(g.view(np.uint32), w.view(np.uint32)), i


@pytest.mark.parametrize("use_fll", [False, True])
def test_fused_mf_batch_ragged_channel_count(gpu, orc, use_fll):
    """70 channels (two full 32-channel CTAs and a partial one), TSC strip and framer behind the fused kernel, three bursts
    with state carried; with the FLL on the call is long enough for the time-chunk pipeline."""
    fs, rs = 10_000_000, 5_000_000
    alpha = float(np.float32(0.4))
    C = 70
    om = orc.QPSKModulator(fs, rs, alpha, 10, True, TSC)
    rng = np.random.default_rng(12)
    n_payload = 300 if use_fll else 64
    rx = []
    for c in range(C):
        s = om.ModulateBytes(rng.integers(0, 256, n_payload, dtype=np.uint8).tobytes(), b"S", b"E")
        a, b = orc.NCO(100e6, fs, 1, seed=31, stream=4 * c), orc.NCO(100e6, fs, 1, seed=31, stream=4 * c + 1)
        rx.append(orc.channel_apply(a, b, 1, s, orc.noise_iq(-40.0, s.size // 2, 31, 4 * c + 2)))
    rx = np.stack(rx)
    gd = gpu.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC, use_fll=use_fll, channels=C)
    ods = [orc.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC, use_fll=use_fll) for _ in range(C)]
    for rep in range(3):
        got = gd.DeModulateBytes(rx, b"S", b"E")
        for c in range(C):
            assert got[c] == ods[c].DeModulateBytes(rx[c], b"S", b"E"), (rep, c)
    assert sum(bool(g) for g in got) > C // 2
