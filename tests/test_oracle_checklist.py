"""CPU: known-answer checklist for the oracle — the behaviours of SURVEY.md §4 items 1-12 that a
"reasonable" re-implementation would silently change.  The reference ships no tests of its own, so
each check is derived by hand from the cited C# lines and pinned here.
"""
import math

import numpy as np
import pytest

TSC = "11001010011101100100100110101100" + "01110100111001011010001101101001"
INV = np.float32(0.7071067811865475)


def _bits(n, seed):
    return "".join(np.random.default_rng(seed).choice(["0", "1"], n))


# 1. Math.Round is banker's rounding; Modulate / FLL use integer division for sps
def test_item1_rounding_and_integer_sps(orc):
    assert orc.RRCFilter.generateCoefficents(6, 0.35, 2500, 1000).size == 6 * 2 + 1      # round(2.5) = 2 (RRC-filter.cs:24)
    assert orc.RRCFilter.generateCoefficents(6, 0.35, 3500, 1000).size == 6 * 4 + 1      # round(3.5) = 4
    assert orc.RRCFilter.generateCoefficents(2.5, 0.35, 4000, 1000).size == 2 * 4 + 1    # span 2.5 -> 2 (:26)
    m = orc.QPSKModulator(3500, 1000, 0.35, 6)                                           # 25 taps (sps 4), Modulate sps = 3
    y = m.Modulate("0011", False)
    assert y.size == 2 * (12 + 2 * 3)                                                    # delay + nDibits * (3500/1000 int)
    nz = np.nonzero(y[0::2])[0]
    assert nz.tolist() == [12, 15]


# 2. RRC singular branches
def test_item2_rrc_singular_branches(orc):
    h = orc.RRCFilter.generateCoefficents(4, 0.25, 8000, 1000)      # 1/(4*beta) = 1 symbol = tap mid +- 8
    mid = (h.size - 1) // 2
    raw0 = 1.0 + 0.25 * (4.0 / math.pi - 1.0)
    raw1 = (0.25 / math.sqrt(2.0)) * ((1.0 + 2.0 / math.pi) * math.sin(math.pi) + (1.0 - 2.0 / math.pi) * math.cos(math.pi))
    assert h[mid + 8] == h[mid - 8]
    assert h[mid + 8] / h[mid] == pytest.approx(raw1 / raw0, rel=1e-12)
    assert np.all(np.isfinite(h))
    assert np.allclose(h, h[::-1], rtol=0, atol=1e-15)


# 3. fftFilter alignment (offset N-1, same length); streaming filter: zero state, natural delay
def test_item3_alignment(orc):
    taps = orc.real_taps_to_iq(orc.RRCFilter.generateCoefficents(6, 0.35, 4000, 1000))   # 25 taps
    n = taps.size // 2
    x = np.zeros(2 * 100, np.float32)
    x[2 * 40] = 1.0
    y = orc.ComplexFIRFilter(taps).fftFilter(x)
    assert y.size == x.size
    assert np.array_equal(y[2 * (40 - (n - 1)):2 * 41:2], taps[0::2])                    # h[0] lands at 40-(N-1)
    ys = orc.ComplexFIRFilter(taps).Filter(x)
    assert np.array_equal(ys[2 * 40:2 * (40 + n):2], taps[0::2])                         # causal: h[0] at 40
    assert not ys[:2 * 40].any()


# 4. streaming dot-product summation order: 8 lanes, then lanes 0..7, then scalar tail; no FMA
def test_item4_lane_order(orc):
    n = 19
    rng = np.random.default_rng(4)
    taps = (rng.standard_normal(2 * n) * 1e3).astype(np.float32)
    x = (rng.standard_normal(2 * n) * 1e-3).astype(np.float32)
    y = orc.ComplexFIRFilter(taps).Filter(x)
    hI, hQ = taps[0::2][::-1], taps[1::2][::-1]
    xI, xQ = x[0::2], x[1::2]
    f = np.float32
    lane_i = [f(0)] * 8
    lane_q = [f(0)] * 8
    for i in range(0, 16, 8):
        for l in range(8):
            lane_i[l] = f(lane_i[l] + f(f(hI[i + l] * xI[i + l]) - f(hQ[i + l] * xQ[i + l])))
            lane_q[l] = f(lane_q[l] + f(f(hI[i + l] * xQ[i + l]) + f(hQ[i + l] * xI[i + l])))
    ai = aq = f(0)
    for l in range(8):
        ai, aq = f(ai + lane_i[l]), f(aq + lane_q[l])
    for i in range(16, n):
        ai = f(ai + f(f(hI[i] * xI[i]) - f(hQ[i] * xQ[i])))
        aq = f(aq + f(f(hI[i] * xQ[i]) + f(hQ[i] * xI[i])))
    assert y[-2] == ai and y[-1] == aq
    assert orc.lib().orc_simd_lanes() == 8 and orc.lib().orc_built_with_avx2() == 1


# 5. Mueller-Muller start-up and limits
def test_item5_mm(orc):
    x = np.arange(2 * 40, dtype=np.float32)
    mm = orc.MuellerMuller(4.0, 0.0, 0.0)
    y = mm.Process(x)
    assert y[0] == x[2] and y[1] == x[3]                           # first symbol: baseIndex 1, mu 0, no TED update
    assert np.array_equal(y[0::2], x[2::8][: y.size // 2])          # kp = ki = 0: advance exactly sps
    big = orc.MuellerMuller(4.0, 1e6, 0.0)                          # huge gain: correction clamps to +-0.1
    z = big.Process(np.tile(np.array([1, -1, -1, 1, 1, 1, -1, -1], np.float32), 50))
    st = big.state
    assert 0.0 <= st["mu"] < 1.0
    one = orc.MuellerMuller(4.0, 0.01, 0.001)
    sig = orc.fill_uniform(3, 0, 0, 2 * 500)
    a = one.Process(sig)
    two = orc.MuellerMuller(4.0, 0.01, 0.001)
    b = np.concatenate([two.Process(sig[:2 * 7]), two.Process(sig[2 * 7:2 * 8]), two.Process(sig[2 * 8:])])
    assert np.array_equal(a, b)                                     # chunk invariant
    assert z.size > 0


# 6. sign decisions: >= 0 -> +1 (zero decides (+1,+1))
def test_item6_sign_of_zero(orc):
    d = orc.QPSKDeModulator(4000, 1000, differentialEncoding=False)
    bits = d.DeModulate(np.zeros(2 * 64, np.float32))
    assert bits and set(bits) == {"1"}


# 7. Costas wraps by one 2*pi step at +-pi; FLL output uses the pre-update phase
def test_item7_wraps(orc):
    c = orc.CostasLoopQpsk(1000.0, 20.0)
    n = np.arange(4000)
    z = np.exp(1j * (0.05 * n))                                     # steady rotation: theta has to keep wrapping
    x = np.empty(2 * n.size, np.float32)
    x[0::2], x[1::2] = z.real, z.imag
    wraps = 0
    prev = 0.0
    for k in range(0, x.size, 400):
        c.Process(x[k:k + 400])
        th, _ = c.GetState()
        assert -math.pi - 0.2 <= th <= math.pi + 0.2                # one +-2*pi step keeps it next to [-pi, pi]
        wraps += th < prev
        prev = th
    assert wraps >= 2
    f = orc.FLLBandEdgeFilter(4.0, 0.35, 40, 0.01)
    f.state = (np.float32(0.5), np.float32(0.0))
    y = f.Process(np.array([1.0, 0.0], np.float32))
    assert y[0] == np.float32(math.cos(np.float32(0.5))) and y[1] == np.float32(math.sin(np.float32(0.5)))
    g = orc.FLLBandEdgeFilter(4.0, 0.35, 40, 0.01)
    g.state = (np.float32(7.0), np.float32(100.0))
    g.Process(np.array([0.0, 0.0], np.float32))
    ph, fr = g.state
    assert fr == np.float32(2.0 * np.float32(math.pi)) * np.float32(0.5)            # clamp to 2*pi*2/sps (:58,191-195)
    # phase += freq happens before the clamp (:124-125), then IEEERemainder brings it into [-pi, pi] (:185-189)
    want = np.float32(math.remainder(float(np.float32(7.0) + np.float32(100.0)), float(np.float32(2.0) * np.float32(math.pi))))
    assert ph == want


# 8. FLL taps: fp32 design, integer mid (asymmetric for even size), upper = conj(lower)
def test_item8_fll_taps(orc):
    lo, up = orc.FLLBandEdgeFilter(2.0, 0.4, 40, 1e-4).taps()
    assert lo.size == 80 and np.array_equal(lo[0::2], up[0::2]) and np.array_equal(lo[1::2], -up[1::2])
    mag = np.hypot(lo[0::2], lo[1::2])
    assert int(np.argmax(mag)) == 19                                 # mid = (40-1)/2 = 19, not 19.5
    assert not np.allclose(mag, mag[::-1])


# 9. differential reference / maps
def test_item9_differential(orc):
    m = orc.QPSKModulator(1000, 1000, 0.35, 0)                       # span 0 -> 1 tap, delay 0, sps 1: raw symbols
    s = m.Modulate("0001" + "11" + "10", False).reshape(-1, 2)
    ref = np.array([INV, INV], np.float32)                           # reset to (1,1)/sqrt2 every call (:126)
    assert np.array_equal(s[0], ref)                                 # 00 -> *1
    assert np.array_equal(s[1], [-INV, INV])                         # 01 -> *j
    assert np.array_equal(s[2], [INV, -INV])                         # 11 -> *-1
    assert np.array_equal(s[3], [-INV, -INV])                        # 10 -> *-j
    assert np.array_equal(m.Modulate("00", False).reshape(-1, 2)[0], ref)
    a = orc.QPSKModulator(1000, 1000, 0.35, 0, False).Modulate("00011110", False).reshape(-1, 2)
    assert np.array_equal(a, np.array([[-INV, -INV], [-INV, INV], [INV, INV], [INV, -INV]], np.float32))


# 10. TSC strip is an exact match inside this call's bits; a miss returns ""
def test_item10_tsc(orc):
    d = orc.QPSKDeModulator(4000, 1000, tsc=TSC)
    assert d.DeModulate(orc.fill_uniform(1, 0, 0, 2 * 400) * 0.1) == ""
    m = orc.QPSKModulator(4000, 1000, 0.35, 6, True, "  ")          # whitespace TSC = none (:27)
    assert m.Modulate("0110").size == orc.QPSKModulator(4000, 1000, 0.35, 6).Modulate("0110").size


# 11. marker hunt: bit offsets outermost, lowest offset wins; carry of 8*len+7 bits
def test_item11_marker_hunt(orc):
    BP = orc.BitPacker
    assert BP.BitsToBytes("0" * 7, 0) == b"" and BP.BitsToBytes("1" * 9, 2) == b""      # usable < 8 -> empty
    assert BP.BitsToBytes("000000011", 1) == b"\x03"
    assert BP.IndexOf(b"abcabc", b"ca") == 2 and BP.IndexOf(b"abc", b"") == 0 and BP.IndexOf(b"ab", b"abc") == -1


# 12. odd trailing bit dropped; odd-length float input throws
def test_item12_odd_inputs(orc):
    m = orc.QPSKModulator(4000, 1000, 0.35, 6)
    assert np.array_equal(m.Modulate("011"), m.Modulate("01"))
    with pytest.raises(orc.ArgumentException):
        orc.QPSKDeModulator(4000, 1000).DeModulate(np.zeros(5, np.float32))
    with pytest.raises(orc.ArgumentException):
        orc.ComplexFIRFilter(np.ones(4, np.float32)).Filter(np.zeros(3, np.float32))


def test_datalevel_roundtrip_config1(orc):
    """BASELINE.json configs[0]: testAtDataLevel — the oracle reproduces the reference's observable
    behaviour (the payload comes back once the loops have pulled in)."""
    fs = 10_000_000
    a = float(np.float32(0.4))
    mod = orc.QPSKModulator(fs, fs // 2, a, 10, tsc=TSC)
    dem = orc.QPSKDeModulator(fs, fs // 2, a, 10, tsc=TSC)
    tx = orc.NCO(100e6, fs, 1, seed=7, stream=0)
    rx = orc.NCO(100e6, fs, 1, seed=7, stream=1)
    text = "The Quick Brown fox jump yes yes man good!"
    got = []
    for _ in range(6):
        s = mod.ModulateTextUtf8(text, "MESSAGE_START", "MESSAGE_STOP")
        assert s.size == 2 * 620
        got.append(dem.DeModulateTextUtf8(orc.channel_apply(tx, rx, 0, s), "MESSAGE_START", "MESSAGE_STOP"))
    assert got[0] == "" and all(g == text for g in got[1:])


def test_nco_and_noise_statistics(orc):
    fs = 10_000_000
    n = orc.NCO(100e6, fs, 1.0, 0.25, seed=5, stream=0)
    z = n.GenerateBlock(30000).reshape(-1, 2)
    assert np.allclose(np.hypot(z[:, 0], z[:, 1]), 1.0, atol=1e-12)
    ph = np.unwrap(np.arctan2(z[:, 1], z[:, 0]))
    f_est = np.diff(ph).mean() * fs / (2 * math.pi)
    f_alias = 100e6 % fs if (100e6 % fs) < fs / 2 else (100e6 % fs) - fs
    assert abs(f_est - f_alias) <= 100e6 * 1.001e-6 + 1.0            # within +-1 ppm (+ drift bound)
    stable = orc.NCO(1e6, fs, 0.0, 0.0, seed=5, stream=0).GenerateBlock(4).reshape(-1, 2)
    assert stable[0, 0] == math.cos(2 * math.pi * 1e6 / fs)           # first sample already advanced (:73-78)
    nz = orc.noise_iq(-20.0, 200000, 9, 2)
    assert abs(nz[0::2].std() - 0.1) < 2e-3 and abs(nz[1::2].std() - 0.1) < 2e-3 and np.abs(nz).max() <= 1.0
    assert np.array_equal(orc.noise_iq(-20.0, 100, 9, 2, first_sample=50)[:100], nz[100:200])


def test_save_as_cs16_known_answers():
    """SaveAsCs16 (MS/Models/HelperFunctions.cs:75-106), hand-derived: normalise by the largest |component|, scale by
    short.MaxValue, truncate toward zero; an all-zero buffer is normalised by 1."""
    from oracle import np_twin
    out, m = np_twin.save_as_cs16(np.array([1.0, -1.0, 0.5, -0.5, 0.25, 0.0], np.float32))
    assert m == 1.0
    assert out.tolist() == [32767, -32767, 16383, -16383, 8191, 0]      # 16383.5 -> 16383, -16383.5 -> -16383
    out, m = np_twin.save_as_cs16(np.array([2.0, -4.0], np.float32))
    assert m == 4.0 and out.tolist() == [16383, -32767]
    out, m = np_twin.save_as_cs16(np.zeros(4, np.float32))
    assert m == 0.0 and out.tolist() == [0, 0, 0, 0]
    with pytest.raises(ValueError):
        np_twin.save_as_cs16(np.zeros(0, np.float32))


def test_decimate_oracle_is_the_subsampled_streaming_filter(orc):
    """Row N1: y_dec[m] = y[skip + m*D] of ComplexFIRFilter.Filter (FIRFilter.cs:80-91), whatever the chunking of the
    full-rate filter underneath (its delay line is carried, :50-52)."""
    import numpy as np
    taps = orc.real_taps_to_iq(orc.RRCFilter.generateCoefficents(8, 0.35, 4000, 1000))
    x = orc.fill_uniform(3, 0, 0, 2 * 1001)
    y = orc.ComplexFIRFilter(taps).Filter(x)
    for D, skip in ((2, 0), (4, 0), (4, 3), (16, 5), (1, 0)):
        d = orc.decimate(y, D, skip)
        assert d.size == 2 * len(range(skip, 1001, D))
        assert np.array_equal(d.reshape(-1, 2), y.reshape(-1, 2)[skip::D])
    f = orc.ComplexFIRFilter(taps)
    y2 = np.concatenate([f.Filter(x[:2 * 333]), f.Filter(x[2 * 333:])])
    assert np.array_equal(orc.decimate(y2, 4), orc.decimate(y, 4))
