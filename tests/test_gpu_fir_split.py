"""GPU parity of fir_split2_kernel (QPSK_FIR_SPLIT; what QPSK_FIR_FAST picks for long real-tap filters): the 2-parallel
fast-FIR split against the CPU oracle and the fp64 reference, over every tail length of its seven-sub-tap blocks, tile
seams, streaming cuts, batches and the stateless fftFilter geometry."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5          # north_star: max |err| <= 1e-5 * max|y|


def _real_taps(n, seed):
    rng = np.random.default_rng(seed)
    t = np.zeros(2 * n, dtype=np.float32)
    t[0::2] = (rng.standard_normal(n) / np.sqrt(n)).astype(np.float32)
    return t


def _x(orc, n_complex, seed=1, stream=0):
    return orc.fill_uniform(seed, stream, 0, 2 * n_complex)


def _split(gpu, taps, channels=1):
    f = gpu.ComplexFIRFilter(taps, channels=channels) if channels > 1 else gpu.ComplexFIRFilter(taps)
    f.set_mode(gpu.FIR_SPLIT)
    return f


@pytest.mark.parametrize("ntaps", list(range(14, 29)) + [33, 64, 65, 96, 97, 129, 200, 257, 513, 769])
def test_split_matches_oracle_every_tail(gpu, orc, ntaps):
    """Gh = ceil(ntaps / 2) sub-taps: 14..28 taps walk through every remainder of the seven-sub-tap blocks."""
    taps = _real_taps(ntaps, ntaps)
    x = _x(orc, 6000, seed=2)
    want = orc.ComplexFIRFilter(taps).Filter(x)
    f = _split(gpu, taps)
    got = f.Filter(x)
    assert "split2" in f.last_kernel(), f.last_kernel()
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= REL_TOL * scale
    assert np.abs(got - orc.fir_filter_f64(taps, x)).max() <= REL_TOL * scale


@pytest.mark.parametrize("L", [1, 2, 9, 2559, 2560, 2561, 5121, 40003])
def test_split_ragged_lengths_and_tile_seams(gpu, orc, L):
    taps = _real_taps(129, 7)
    x = _x(orc, L, seed=3)
    want = orc.ComplexFIRFilter(taps).Filter(x)
    got = _split(gpu, taps).Filter(x)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= REL_TOL * max(np.abs(want).max(), 1e-30)


def test_split_streaming_cuts_carry_the_delay_line(gpu, orc):
    taps = _real_taps(257, 11)
    x = _x(orc, 30000, seed=4)
    want = orc.ComplexFIRFilter(taps).Filter(x)
    f = _split(gpu, taps)
    cuts = [0, 2, 600, 602, 5000, 5122, 17000, 2 * 30000]
    got = np.concatenate([f.Filter(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
    assert np.abs(got - want).max() <= REL_TOL * np.abs(want).max()
    # no bit-level chunk invariance in this mode (unlike the FMA and EXACT kernels): the recombination y[2m+1] = C - A' - B
    # cancels a term that contains the NEXT sample (zero at the end of a chunk, real data inside a longer call), exactly in
    # real arithmetic, to rounding in fp32 — and a cut at an odd index swaps the roles of the phases.  Tolerance only.
    one = _split(gpu, taps).Filter(x)
    assert np.abs(one - got).max() <= REL_TOL * np.abs(want).max()


def test_split_batch_of_channels(gpu, orc):
    taps = _real_taps(97, 5)
    C, L = 5, 7002          # even row stride: the TMA kernels need 16-byte aligned rows
    x = np.stack([_x(orc, L, seed=6, stream=c) for c in range(C)])
    f = _split(gpu, taps, channels=C)
    got = f.Filter(x)
    assert "split2" in f.last_kernel()
    for c in range(C):
        want = orc.ComplexFIRFilter(taps).Filter(x[c])
        assert np.abs(got[c] - want).max() <= REL_TOL * np.abs(want).max()


def test_split_falls_back_when_the_ring_does_not_fit(gpu, orc):
    """1025 taps: two stages of tile + halo, the output slice and the S planes exceed the shared memory of a CTA — the FMA
    kernel takes the call, results as ever."""
    taps = _real_taps(1025, 13)
    x = _x(orc, 6000, seed=2)
    f = _split(gpu, taps)
    got = f.Filter(x)
    assert "fir_tma_kernel" in f.last_kernel(), f.last_kernel()
    want = orc.ComplexFIRFilter(taps).Filter(x)
    assert np.abs(got - want).max() <= REL_TOL * np.abs(want).max()


def test_split_fft_filter_geometry(gpu, orc):
    """fftFilter (FIRFilter.cs:96-142) = the stateless alignment of the same kernel."""
    taps = _real_taps(129, 9)
    x = _x(orc, 9000, seed=8)
    want = orc.ComplexFIRFilter(taps).fftFilter(x)
    got = _split(gpu, taps).fftFilter(x)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= REL_TOL * np.abs(want).max()


def test_fast_mode_dispatch_and_fma_mode(gpu, orc):
    """QPSK_FIR_FAST picks the split kernel for long real-tap filters only; QPSK_FIR_FMA never does; complex taps and
    QPSK_FIR_EXACT never do."""
    x = _x(orc, 4000, seed=9)
    long_real, short_real = _real_taps(257, 1), _real_taps(33, 2)
    f = gpu.ComplexFIRFilter(long_real)
    a = f.Filter(x)
    assert "split2" in f.last_kernel()
    g = gpu.ComplexFIRFilter(long_real)
    g.set_mode(gpu.FIR_FMA)
    b = g.Filter(x)
    assert "fir_tma_kernel" in g.last_kernel()
    assert np.abs(a - b).max() <= REL_TOL * np.abs(b).max()
    h = gpu.ComplexFIRFilter(short_real)
    h.Filter(x)
    assert "fir_tma_kernel" in h.last_kernel()
    rng = np.random.default_rng(3)
    ct = (rng.standard_normal(2 * 257) / 16).astype(np.float32)
    c = gpu.ComplexFIRFilter(ct)
    c.set_mode(gpu.FIR_SPLIT)
    c.Filter(x)
    assert "split2" not in c.last_kernel()
    e = gpu.ComplexFIRFilter(long_real)
    e.set_mode(gpu.FIR_EXACT)
    got = e.Filter(x)
    assert "exact" in e.last_kernel()
    want = orc.ComplexFIRFilter(long_real).Filter(x)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
