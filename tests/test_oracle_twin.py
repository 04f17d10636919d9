"""CPU: the C++ oracle against its independently written numpy twin (oracle/np_twin.py).

Both restate the same C# sources; neither can be checked against the reference itself (no .NET in
this image, no golden vectors upstream: parity unpinned).  Agreement of two separate transcriptions
is the guard against transcription mistakes.  fp32 paths and the fp64 loops must agree bit for bit
(same libm); fftFilter goes through a real FFT in the twin and a direct fp64 convolution in the
oracle, so it agrees to fp64 round-off (<= 1 fp32 ulp after the cast).
"""
import numpy as np
import pytest

from oracle import np_twin as T


def _bits(n, seed):
    return "".join(np.random.default_rng(seed).choice(["0", "1"], n))


def _biteq(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def _impaired(orc, nbits=400, sps=4, span=6, cfo=0.02, seed=0):
    x = orc.QPSKModulator(sps * 1000, 1000, 0.35, span).Modulate(_bits(nbits, seed))
    z = (x[0::2] + 1j * x[1::2]) * np.exp(1j * (0.3 + cfo * np.arange(x.size // 2)))
    y = np.empty_like(x)
    y[0::2], y[1::2] = z.real, z.imag
    return y


@pytest.mark.parametrize("span,beta,fs,rs", [
    (10, float(np.float32(0.4)), 10_000_000, 5_000_000), (16, 0.35, 4000, 1000), (4, 0.25, 8000, 1000),
    (6, 0.5, 2000, 1000), (6, 0.35, 2500, 1000), (11, 0.9, 10_000_000, 333333), (5, 0.5, 3000, 1000), (2.5, 0.35, 3500, 1000)])
def test_rrc_taps(orc, span, beta, fs, rs):
    a = orc.RRCFilter.generateCoefficents(span, beta, fs, rs)
    b = T.rrc_taps(span, beta, fs, rs)
    assert np.array_equal(a, b)
    assert abs(np.sum(a * a) - 1.0) < 1e-12


@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 21, 40])
def test_fir_streaming_and_fft(orc, n):
    rng = np.random.default_rng(n)
    taps = rng.standard_normal(2 * n).astype(np.float32)
    x = orc.fill_uniform(1, 0, 0, 2 * 257)
    fo, ft = orc.ComplexFIRFilter(taps), T.ComplexFIRFilter(taps)
    assert _biteq(fo.Filter(x[:100]), ft.Filter(x[:100]))
    assert _biteq(fo.Filter(x[100:]), ft.Filter(x[100:]))          # state carried
    a, b = fo.fftFilter(x), ft.fftFilter(x)
    assert np.abs(a - b).max() <= 2e-7 * np.abs(a).max()


@pytest.mark.parametrize("diff,tsc", [(True, "1100101001110110"), (False, "101"), (True, None)])
def test_modulator(orc, diff, tsc):
    bits = _bits(301, 3)
    a = orc.QPSKModulator(4000, 1000, 0.35, 6, diff, tsc).Modulate(bits)
    b = T.QPSKModulator(4000, 1000, 0.35, 6, diff, tsc).Modulate(bits)
    assert a.shape == b.shape and np.abs(a - b).max() <= 2e-7 * np.abs(a).max()
    a0 = orc.QPSKModulator(4000, 1000, 0.35, 6, diff, tsc).Modulate(bits, False)
    b0 = T.QPSKModulator(4000, 1000, 0.35, 6, diff, tsc).Modulate(bits, False)
    assert _biteq(a0, b0)
    ab = orc.QPSKModulator(4000, 1000, 0.35, 6, diff, tsc).ModulateBytes(b"hello", b"<", b">")
    bb = T.QPSKModulator(4000, 1000, 0.35, 6, diff, tsc).ModulateBytes(b"hello", b"<", b">")
    assert ab.shape == bb.shape and np.abs(ab - bb).max() <= 2e-7 * np.abs(ab).max()


@pytest.mark.parametrize("sps,rolloff,size,bw", [(4.0, 0.35, 40, 0.01), (30.0, 0.9, 10, 0.1), (2.0, 0.4, 13, 0.05)])
def test_fll(orc, sps, rolloff, size, bw):
    y = _impaired(orc, 300, sps=int(sps), seed=int(sps))
    fo, ft = orc.FLLBandEdgeFilter(sps, rolloff, size, bw), T.FLLBandEdgeFilter(sps, rolloff, size, bw)
    lo, up = fo.taps()
    assert np.array_equal(lo, ft.lower_taps) and np.array_equal(up, ft.upper_taps)
    assert _biteq(fo.Process(y), ft.Process(y))
    assert fo.state == (float(ft.phase), float(ft.freq))


def test_mm_costas_and_demod(orc):
    y = _impaired(orc, 500, seed=5)
    mf = orc.ComplexFIRFilter(orc.real_taps_to_iq(orc.RRCFilter.generateCoefficents(6, 0.35, 4000, 1000))).Filter(y)
    kp, ki = orc.mm_gains_from_bw(0.01)
    mo, mt = orc.MuellerMuller(4.0, kp, ki), T.MuellerMuller(4.0, kp, ki)
    a = np.concatenate([mo.Process(mf[:600]), mo.Process(mf[600:])])
    b = np.concatenate([mt.Process(mf[:600]), mt.Process(mf[600:])])
    assert _biteq(a, b)
    assert _biteq(orc.MuellerMuller(4.0, kp, ki).Process(mf, 5), T.MuellerMuller(4.0, kp, ki).Process(mf, 5))
    assert _biteq(orc.CostasLoopQpsk(1000, 100).Process(a), T.CostasLoopQpsk(1000, 100).Process(b))
    for fll in (False, True):
        for tsc in (None, "110010100111"):
            kw = dict(SymbolSyncBandwith=0.002, CostasLoopBandwith=120.0, tsc=tsc, use_fll=fll)
            od = orc.QPSKDeModulator(4000, 1000, 0.35, 6, **kw)
            td = T.QPSKDeModulator(4000, 1000, 0.35, 6, 0.002, 120.0, tsc=tsc, use_fll=fll)
            assert od.DeModulate(y[:1000]) + od.DeModulate(y[1000:]) == td.DeModulate(y[:1000]) + td.DeModulate(y[1000:])


def test_framer(orc):
    sm, em = b"\xa5ST", b"EN\x5a"
    mod = orc.QPSKModulator(4000, 1000, 0.35, 6)
    x = np.concatenate([mod.ModulateBytes(p, sm, em) for p in (b"first payload", bytes(range(60)), b"x")])
    od = orc.QPSKDeModulator(4000, 1000, 0.35, 6, SymbolSyncBandwith=0.002)
    td = T.QPSKDeModulator(4000, 1000, 0.35, 6, 0.002)
    outs = []
    for a in range(0, x.size, 2 * 211):
        w, g = od.DeModulateBytes(x[a:a + 422], sm, em), td.DeModulateBytes(x[a:a + 422], sm, em)
        assert w == g
        assert od.in_frame == td.in_frame
        outs.append(w)
    assert any(outs)


def test_framer_on_bits(orc):
    """The framer alone (QPSKDeModulator.cs:179-259) on random bit chunks with frames at every bit offset."""
    sm, em = b"\xa5ST", b"EN\x5a"
    rng = np.random.default_rng(11)
    bits_of = lambda b: "".join(format(v, "08b") for v in b)
    od = orc.QPSKDeModulator(4000, 1000)
    td = T.QPSKDeModulator(4000, 1000)
    frames = 0
    for trial in range(400):
        n = int(rng.integers(0, 80))
        if rng.random() < 0.3:
            pad = "".join(rng.choice(["0", "1"], int(rng.integers(0, 9))))
            chunk = (pad + bits_of(sm + rng.integers(0, 256, int(rng.integers(0, 6)), dtype=np.uint8).tobytes() + em))[: n + 30]
        else:
            chunk = "".join(rng.choice(["0", "1"], n))
        w, g = od.FrameBits(chunk, sm, em, cap=4096), td.FrameBits(chunk, sm, em)
        assert w == g, trial
        assert od.in_frame == td.in_frame
        frames += bool(w) or (od.in_frame is False and chunk.endswith(bits_of(em)))
    assert frames > 5


def test_bitpacker_mirror(orc):
    """The product's host-side BitPacker mirror (HelperFunctions.cs:11-70) against the oracle's."""
    import qpsk_modulator_demodulator_b200.modem as M
    rng = np.random.default_rng(4)
    for n in (0, 1, 2, 17, 300):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        s = orc.BitPacker.BytesToBitString(data)
        assert M.BitPacker.BytesToBitString(data) == s
        for off in range(8):
            assert M.BitPacker.BitsToBytes(s + "101", off) == orc.BitPacker.BitsToBytes(s + "101", off)
    assert M.BitPacker.BitsToBytes("1010101", 0) == b""                      # fewer than 8 usable bits (:38-39)
    with pytest.raises(ValueError):
        M.BitPacker.BitsToBytes("10101010", 8)                                # :35
    with pytest.raises(M.FormatException):
        M.BitPacker.BitsToBytes("1010x010", 0)                                # :50
    for hay, nee in ((b"abcabc", b"ca"), (b"abc", b""), (b"ab", b"abc"), (b"", b""), (b"xyz", b"q")):
        assert M.BitPacker.IndexOf(hay, nee) == orc.BitPacker.IndexOf(hay, nee)


def test_text_marker_defaults(orc):
    """ModulateTextUtf8 / DeModulateTextUtf8 default to STX / ETX (QPSKModulator.cs:76-77, QPSKDeModulator.cs:264-265)."""
    import inspect
    import qpsk_modulator_demodulator_b200.modem as M
    for cls in (orc.QPSKModulator, M.QPSKModulator):
        p = inspect.signature(cls.ModulateTextUtf8).parameters
        assert (p["startMarker"].default, p["endMarker"].default) == ("\x02", "\x03")
    for cls in (orc.QPSKDeModulator, M.QPSKDeModulator):
        p = inspect.signature(cls.DeModulateTextUtf8).parameters
        assert (p["startMarker"].default, p["endMarker"].default) == ("\x02", "\x03")


def test_bitpacker(orc):
    data = bytes(range(0, 256, 7))
    s = orc.BitPacker.BytesToBitString(data)
    assert s == T.bytes_to_bit_string(data)
    for off in range(8):
        assert orc.BitPacker.BitsToBytes(s, off) == T.bits_to_bytes(s, off)
    assert orc.BitPacker.BitsToBytes(s, 0) == data
