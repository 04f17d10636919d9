"""GPU parity: ComplexFIRFilter (K1/K2) through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# fp32 tolerance for filter outputs (north_star): max |err| <= 1e-5 * max|y| against the
# reference's fp64 Complex path; the FAST kernel differs from the oracle's fp32 path only by
# summation order / FMA contraction.
REL_TOL = 1e-5


def _rand_iq(orc, n_complex, seed=1, stream=0):
    return orc.fill_uniform(seed, stream, 0, 2 * n_complex)


def _rrc_iq(orc, span, sps, alpha=0.35):
    h = orc.RRCFilter.generateCoefficents(span, alpha, sps * 1000, 1000)
    return orc.real_taps_to_iq(h)


@pytest.mark.parametrize("span,sps", [(16, 2), (8, 4), (4, 8), (10, 2), (32, 8), (16, 16), (5, 3), (1, 1)])
@pytest.mark.parametrize("L", [1, 7, 620, 2560, 2561, 40000])
def test_streaming_real_taps_matches_oracle(gpu, orc, span, sps, L):
    taps = _rrc_iq(orc, span, sps)
    x = _rand_iq(orc, L)
    want = orc.ComplexFIRFilter(taps).Filter(x)
    got = gpu.ComplexFIRFilter(taps).Filter(x)
    scale = max(np.abs(want).max(), 1e-30)
    assert np.abs(got - want).max() <= REL_TOL * scale
    ref64 = orc.fir_filter_f64(taps, x)
    assert np.abs(got - ref64).max() <= REL_TOL * scale


@pytest.mark.parametrize("ntaps", [1, 2, 3, 8, 10, 33, 40, 64, 257])
def test_streaming_complex_taps_matches_oracle(gpu, orc, ntaps):
    rng = np.random.default_rng(ntaps)
    taps = (rng.standard_normal(2 * ntaps) / np.sqrt(ntaps)).astype(np.float32)
    x = _rand_iq(orc, 9000, seed=3)
    want = orc.ComplexFIRFilter(taps).Filter(x)
    got = gpu.ComplexFIRFilter(taps).Filter(x)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= REL_TOL * scale


@pytest.mark.parametrize("ntaps,cplx", [(21, False), (33, False), (40, True), (7, True), (257, False)])
def test_exact_mode_is_bit_identical_to_reference_order(gpu, orc, ntaps, cplx):
    """QPSK_FIR_EXACT reproduces ComplexDotWindow's 8-lane summation order (FIRFilter.cs:165-192)."""
    rng = np.random.default_rng(100 + ntaps)
    taps = (rng.standard_normal(2 * ntaps) / np.sqrt(ntaps)).astype(np.float32)
    if not cplx:
        taps[1::2] = 0
    x = _rand_iq(orc, 5000, seed=5)
    want = orc.ComplexFIRFilter(taps).Filter(x)
    f = gpu.ComplexFIRFilter(taps)
    f.set_mode(gpu.FIR_EXACT)
    got = f.Filter(x)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("ntaps", [8, 9, 16, 24, 41, 64, 130])
def test_exact_mode_real_taps_special_values(gpu, orc, ntaps):
    """The exact-order TMA kernel for real taps drops the reference's (+-0 * sample) terms; that must stay bit-neutral
    for zeros of either sign, denormals and exact cancellations, and infinite / NaN samples (where 0 * Inf = NaN in the
    reference) must come out as the reference's NaNs.  Several tiles, chunked calls, odd lengths."""
    rng = np.random.default_rng(7 * ntaps)
    taps = np.zeros(2 * ntaps, np.float32)
    taps[0::2] = (rng.standard_normal(ntaps) / np.sqrt(ntaps)).astype(np.float32)
    taps[2 * (ntaps // 3)] = 0.0                      # a zero tap
    if ntaps > 10:
        taps[2 * 5 + 1] = -0.0                        # an imaginary part of -0 is still "real taps"
    x = _rand_iq(orc, 3001, seed=ntaps)
    x[100:160] = 0.0
    x[200:230] = -0.0
    x[300:310] = np.float32(1e-42)                    # denormals
    x[400] = np.inf
    x[1501] = -np.inf
    x[2200] = np.nan
    x[2600:2604] = [1.0, -1.0, -1.0, 1.0]
    want_f = orc.ComplexFIRFilter(taps)
    got_f = gpu.ComplexFIRFilter(taps)
    got_f.set_mode(gpu.FIR_EXACT)
    for a, b in [(0, 2 * 700), (2 * 700, 2 * 701), (2 * 701, x.size)]:
        want = want_f.Filter(x[a:b])
        got = got_f.Filter(x[a:b])
        wn, gn = np.isnan(want), np.isnan(got)
        assert np.array_equal(wn, gn)
        assert np.array_equal(got.view(np.uint32)[~wn], want.view(np.uint32)[~wn])


@pytest.mark.parametrize("mode", ["fast", "exact"])
def test_chunked_equals_one_shot(gpu, orc, mode):
    """State carried across calls: arbitrary chunking gives the same stream."""
    taps = _rrc_iq(orc, 10, 4)
    x = _rand_iq(orc, 20000, seed=9)
    f1 = gpu.ComplexFIRFilter(taps)
    f2 = gpu.ComplexFIRFilter(taps)
    if mode == "exact":
        f1.set_mode(gpu.FIR_EXACT)
        f2.set_mode(gpu.FIR_EXACT)
    one = f1.Filter(x)
    cuts = [0, 2, 4, 38, 40, 1000, 1002, 6122, 16000, 40000]
    parts = [f2.Filter(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    many = np.concatenate(parts)
    if mode == "exact":
        assert np.array_equal(one.view(np.uint32), many.view(np.uint32))
    else:
        # tile boundaries move with the chunking, the per-output arithmetic does not
        assert np.array_equal(one.view(np.uint32), many.view(np.uint32))
    want = orc.ComplexFIRFilter(taps).Filter(x)
    assert np.abs(one - want).max() <= REL_TOL * np.abs(want).max()


def test_state_roundtrip_and_reset(gpu, orc):
    taps = _rrc_iq(orc, 8, 4)
    x = _rand_iq(orc, 3000, seed=11)
    f = gpu.ComplexFIRFilter(taps)
    f.Filter(x[:2000])
    st = f.get_state()
    assert st.shape == (1, 2 * (taps.size // 2 - 1))
    assert np.array_equal(st[0], x[2000 - st.shape[1]:2000])
    g = gpu.ComplexFIRFilter(taps)
    g.set_state(st)
    assert np.array_equal(f.Filter(x[2000:]), g.Filter(x[2000:]))
    f.reset()
    fresh = gpu.ComplexFIRFilter(taps)
    assert np.array_equal(f.Filter(x[:100]), fresh.Filter(x[:100]))


@pytest.mark.parametrize("span,sps,L", [(10, 2, 620), (6, 4, 1), (6, 4, 5), (10, 30, 9000), (16, 2, 2560), (8, 4, 5121)])
def test_fft_filter_alignment_and_values(gpu, orc, span, sps, L):
    """fftFilter: stateless, same length, offset N-1 (FIRFilter.cs:130-138)."""
    taps = _rrc_iq(orc, span, sps, 0.9)
    x = _rand_iq(orc, L, seed=13)
    want = orc.ComplexFIRFilter(taps).fftFilter(x)
    f = gpu.ComplexFIRFilter(taps)
    got = f.fftFilter(x)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= REL_TOL * max(np.abs(want).max(), 1e-30)
    # stateless: a second call gives the same answer and does not disturb the streaming state
    assert np.array_equal(f.fftFilter(x), got)


def test_fft_filter_complex_taps(gpu, orc):
    rng = np.random.default_rng(5)
    taps = rng.standard_normal(2 * 19).astype(np.float32)
    x = _rand_iq(orc, 3001, seed=17)
    want = orc.ComplexFIRFilter(taps).fftFilter(x)
    got = gpu.ComplexFIRFilter(taps).fftFilter(x)
    assert np.abs(got - want).max() <= REL_TOL * np.abs(want).max()


def test_batch_channels_are_independent(gpu, orc):
    taps = _rrc_iq(orc, 10, 4)
    C, L = 7, 6000
    x = np.stack([_rand_iq(orc, L, seed=21, stream=c) for c in range(C)])
    f = gpu.ComplexFIRFilter(taps, channels=C)
    got1 = f.Filter(x[:, :5000])
    got2 = f.Filter(x[:, 5000:])
    got = np.concatenate([got1, got2], axis=1)
    for c in range(C):
        want = orc.ComplexFIRFilter(taps).Filter(x[c])
        assert np.abs(got[c] - want).max() <= REL_TOL * np.abs(want).max()


def test_exact_mode_batch_channels(gpu, orc):
    """QPSK_FIR_EXACT over several channels with an even row pitch (TMA path) and with an odd one (generic kernel):
    every channel bit-identical to its own single-stream oracle, state carried over two calls."""
    taps = _rrc_iq(orc, 10, 4)
    C = 5
    for L in (3000, 3001):
        x = np.stack([_rand_iq(orc, L, seed=31, stream=c) for c in range(C)])
        f = gpu.ComplexFIRFilter(taps, channels=C)
        f.set_mode(gpu.FIR_EXACT)
        got = np.concatenate([f.Filter(np.ascontiguousarray(x[:, :2 * 1700])), f.Filter(np.ascontiguousarray(x[:, 2 * 1700:]))], axis=1)
        for c in range(C):
            want = orc.ComplexFIRFilter(taps).Filter(x[c])
            assert np.array_equal(got[c].view(np.uint32), want.view(np.uint32)), (L, c)


def test_error_behaviour_matches_reference(gpu, orc):
    Q, O = gpu, orc
    for mod in (Q, O):
        with pytest.raises(mod.ArgumentNullException):
            mod.ComplexFIRFilter(None)
        with pytest.raises(mod.ArgumentException):
            mod.ComplexFIRFilter(np.zeros(3, np.float32))     # odd length (FIRFilter.cs:32)
        with pytest.raises(mod.ArgumentException):
            mod.ComplexFIRFilter(np.zeros(0, np.float32))     # empty (FIRFilter.cs:33)
        f = mod.ComplexFIRFilter(np.ones(4, np.float32))
        with pytest.raises(mod.ArgumentException):
            f.Filter(np.zeros(3, np.float32))                 # odd input (:82)
        with pytest.raises(mod.ArgumentException):
            f.Filter(np.zeros(4, np.float32), out_len=2)      # short output (:83)
        with pytest.raises(mod.ArgumentException):
            f.fftFilter(np.zeros(5, np.float32))              # odd (:99)
        assert f.fftFilter(np.zeros(0, np.float32)).size == 0  # empty -> empty (:100)
        assert f.Filter(np.zeros(0, np.float32)).size == 0


def test_device_resident_large_stream_linearity(gpu, orc):
    """Full-size property check (no oracle at this size): FIR is linear and shift-invariant."""
    import torch
    taps = _rrc_iq(orc, 16, 2)
    n = 1 << 22
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    s = ts.cuda_stream     # non-zero: a NULL stream means "the handle's own stream" in the C ABI
    x = torch.empty(2 * n, dtype=torch.float32, device="cuda")
    gpu.fill_uniform_dev(1, 0, 0, 2 * n, x.data_ptr(), s)
    y = torch.empty_like(x)
    f = gpu.ComplexFIRFilter(taps)
    f.filter_dev(x.data_ptr(), y.data_ptr(), 2 * n, stream=s)
    x2 = (2.0 * x).contiguous()
    y2 = torch.empty_like(x)
    g = gpu.ComplexFIRFilter(taps)
    g.filter_dev(x2.data_ptr(), y2.data_ptr(), 2 * n, stream=s)
    torch.cuda.synchronize()
    assert torch.equal(y2, 2.0 * y)          # scaling by 2 is exact in binary fp
    # spot-check a window deep inside the stream against the oracle
    lo = (n // 2) * 2
    seg = x[lo - 2 * 64: lo + 2 * 1000].cpu().numpy()
    want = orc.ComplexFIRFilter(taps).Filter(seg)[2 * 64:]
    got = y[lo: lo + 2 * 1000].cpu().numpy()
    assert np.abs(got - want).max() <= REL_TOL * np.abs(want).max()


@pytest.mark.parametrize("kind", ["pageable", "registered", "pinned"])
def test_host_pipeline_long_stream_equals_oracle(gpu, orc, kind):
    """Host-pointer calls longer than one pipeline chunk (2^21 samples): the chunked H2D/kernel/D2H pipeline
    must give the one-shot stream, for the streaming form (delay line carried across chunks) and for the
    stateless fftFilter form (N-1 look-ahead samples staged with every chunk)."""
    taps = _rrc_iq(orc, 10, 2)
    L = 2 * (1 << 21) + 12345
    x0 = _rand_iq(orc, L, seed=31)
    holders = []
    if kind == "pinned":
        bi, bo = gpu.PinnedBuffer(2 * L), gpu.PinnedBuffer(2 * L)
        holders += [bi, bo]
        x, y = bi.array, bo.array
        x[:] = x0
    else:
        x, y = x0.copy(), np.empty(2 * L, np.float32)
        if kind == "registered":
            holders += [gpu.RegisteredArray(x), gpu.RegisteredArray(y)]
    f = gpu.ComplexFIRFilter(taps)
    f.Filter(x, y)
    want = orc.ComplexFIRFilter(taps).Filter(x0)
    assert np.abs(y - want).max() <= REL_TOL * np.abs(want).max()
    # chunk seams: bit-identical to a device-resident one-shot run of the same kernel
    import torch
    dx = torch.from_numpy(x0).cuda()
    dy = torch.empty_like(dx)
    gpu.ComplexFIRFilter(taps).filter_dev(dx.data_ptr(), dy.data_ptr(), 2 * L)   # one launch over the whole stream
    torch.cuda.synchronize()
    assert np.array_equal(dy.cpu().numpy().view(np.uint32), np.asarray(y).view(np.uint32))
    want_s = orc.ComplexFIRFilter(taps).fftFilter(x0)
    got_s = f.fftFilter(x)
    assert got_s.shape == want_s.shape
    assert np.abs(got_s - want_s).max() <= REL_TOL * np.abs(want_s).max()
    seam = 2 * (1 << 21)
    for k in (1, 2):
        w = slice(k * seam - 200, k * seam + 200)
        assert np.abs(got_s[w] - want_s[w]).max() <= REL_TOL * np.abs(want_s).max()
    for h in holders:
        (h.free if hasattr(h, "free") else h.release)()


def test_full_size_sweep_properties(gpu, orc):
    """BASELINE.json configs[1] at full size (2^28 cf32 samples, 2 GiB in + 2 GiB out): no oracle at this size, so
    check size-independent properties of every tap count of the sweep — impulse response = the taps,
    exact homogeneity under scaling by 2, and oracle agreement on windows at the start, at tile seams deep in
    the stream and at the very end."""
    import torch
    n = 1 << 28
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    s = ts.cuda_stream
    x = torch.empty(2 * n, dtype=torch.float32, device="cuda")
    y = torch.empty(2 * n, dtype=torch.float32, device="cuda")
    gpu.fill_uniform_dev(1, 0, 0, 2 * n, x.data_ptr(), s)
    torch.cuda.synchronize()
    # unit impulses far apart (complex 1 and complex j) -> the taps come out at those offsets
    imp = [(12345, 0), (n // 2 + 1, 1), (n - 300, 0)]
    saved = []
    for pos, comp in imp:
        saved.append(x[2 * pos - 2 * 600: 2 * pos + 2 * 600].clone())
        x[2 * pos - 2 * 600: 2 * pos + 2 * 600] = 0
        x[2 * pos + comp] = 1.0
    for span, sps in [(16, 2), (16, 4), (16, 8), (16, 16)]:
        taps = _rrc_iq(orc, span, sps)
        nt = taps.size // 2
        f = gpu.ComplexFIRFilter(taps)
        f.filter_dev(x.data_ptr(), y.data_ptr(), 2 * n, stream=s)
        torch.cuda.synchronize()
        for pos, comp in imp:
            m = min(nt, n - pos)
            got = y[2 * pos: 2 * (pos + m)].cpu().numpy().reshape(-1, 2)
            if "split2" in f.last_kernel():
                # the split kernel returns odd-phase taps as (h0 + h1) - h0: right to a rounding of the pair sum, not exact
                assert np.abs(got[:, comp] - taps[0:2 * m:2]).max() <= 2e-7 * np.abs(taps).max(), (nt, pos)
                fm = gpu.ComplexFIRFilter(taps)
                fm.set_mode(gpu.FIR_FMA)              # the tap-sequential kernel: the taps themselves, bit for bit
                lo_w = pos - nt
                yw = torch.empty(2 * (m + nt), dtype=torch.float32, device="cuda")
                fm.filter_dev(x[2 * lo_w:].data_ptr(), yw.data_ptr(), 2 * (m + nt), stream=s)
                torch.cuda.synchronize()
                gw = yw[2 * nt:].cpu().numpy().reshape(-1, 2)
                assert np.array_equal(gw[:, comp], taps[0:2 * m:2]), (nt, pos)
                assert not gw[:, 1 - comp].any()
            else:
                assert np.array_equal(got[:, comp], taps[0:2 * m:2]), (nt, pos)
            assert not got[:, 1 - comp].any()
        # windows against the oracle (with enough lead-in to fill the delay line)
        for lo in (0, 2560 * 7 - 100, n // 3, n - 5000):
            a = max(lo - nt, 0)
            seg = x[2 * a: 2 * min(lo + 3000, n)].cpu().numpy()
            want = orc.ComplexFIRFilter(taps).Filter(seg)[2 * (lo - a):]
            got = y[2 * lo: 2 * lo + want.size].cpu().numpy()
            if a > 0:
                # the oracle started from a zero delay line at `a`; the first nt-1 outputs after `a` differ, lo is past them
                assert lo - a >= nt - 1
            assert np.abs(got - want).max() <= REL_TOL * max(np.abs(want).max(), 1e-30), (nt, lo)
        # homogeneity on a strided sample of the output: filter(2x) == 2 filter(x) exactly
        chk = y[:: 4099].clone()
        x.mul_(2.0)
        f2 = gpu.ComplexFIRFilter(taps)
        f2.filter_dev(x.data_ptr(), y.data_ptr(), 2 * n, stream=s)
        torch.cuda.synchronize()
        assert torch.equal(y[:: 4099], 2.0 * chk)
        x.mul_(0.5)
    del x, y
    torch.cuda.empty_cache()
