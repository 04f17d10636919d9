"""CPU oracle bindings — TEST INFRASTRUCTURE (see qpsk_oracle.cpp header; parity unpinned).

ctypes wrappers over oracle/liboracle.so exposing the reference's class surface
(RRCFilter, ComplexFIRFilter, FLLBandEdgeFilter, MuellerMuller, CostasLoopQpsk,
QPSKModulator, QPSKDeModulator, NCO) so parity tests read like calls on the C# classes.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

ERR_NULL, ERR_ARG, ERR_RANGE, ERR_CAPACITY = -1, -2, -3, -6


class ArgumentNullException(ValueError):
    pass


class ArgumentException(ValueError):
    pass


class ArgumentOutOfRangeException(ValueError):
    pass


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "qpsk_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_vpp = C.POINTER(C.c_void_p)


def _declare(L):
    i, i64, u64, d, f, vp, cp = C.c_int, C.c_int64, C.c_uint64, C.c_double, C.c_float, C.c_void_p, C.c_char_p
    ip, i64p = C.POINTER(C.c_int), C.POINTER(C.c_int64)
    sig = {
        "orc_simd_lanes": (i, []),
        "orc_built_with_avx2": (i, []),
        "orc_rng_u64": (u64, [u64, u64, u64]),
        "orc_rng_double": (d, [u64, u64, u64]),
        "orc_fill_uniform": (None, [u64, u64, i64, i64, _f32p]),
        "orc_fill_bytes": (None, [u64, u64, i64, i64, _u8p]),
        "orc_rrc_taps": (i, [d, d, i, i, _f64p, i, ip]),
        "orc_fir_create": (i, [_f32p, i, _vpp]),
        "orc_fir_destroy": (None, [vp]),
        "orc_fir_reset": (None, [vp]),
        "orc_fir_filter": (i, [vp, _f32p, _f32p, i64, i64]),
        "orc_fir_fft_filter": (i, [vp, _f32p, _f32p, i64]),
        "orc_convolve_f64": (None, [_f64p, i64, _f64p, i64, _f64p]),
        "orc_fir_filter_f64": (None, [_f32p, i, _f32p, _f64p, i64]),
        "orc_fir_filter_mt": (i, [_f32p, i, _f32p, _f32p, i64, i]),
        "orc_fll_create": (i, [f, f, i, f, _vpp]),
        "orc_fll_destroy": (None, [vp]),
        "orc_fll_taps": (i, [vp, _f32p, _f32p]),
        "orc_fll_process": (i, [vp, _f32p, _f32p, i64, i64]),
        "orc_fll_get_state": (None, [vp, _f32p, _f32p]),
        "orc_fll_set_state": (None, [vp, f, f]),
        "orc_mm_create": (i, [d, d, d, _vpp]),
        "orc_mm_destroy": (None, [vp]),
        "orc_mm_process": (i, [vp, _f32p, i64, _f32p, i64, ip]),
        "orc_mm_get_state": (None, [vp, ip, _f64p, _f64p, ip]),
        "orc_mm_gains_from_bw": (None, [d, _f64p, _f64p]),
        "orc_costas_create": (i, [d, d, d, _vpp]),
        "orc_costas_destroy": (None, [vp]),
        "orc_costas_process": (i, [vp, _f32p, _f32p, i64, i64]),
        "orc_costas_get_state": (None, [vp, _f64p, _f64p]),
        "orc_costas_gains": (None, [vp, _f64p, _f64p]),
        "orc_bytes_to_bits": (None, [_u8p, i64, cp]),
        "orc_bits_to_bytes": (i64, [cp, i64, i, _u8p, i64]),
        "orc_index_of": (i64, [_u8p, i64, _u8p, i64]),
        "orc_mod_create": (i, [i, i, d, i, i, cp, _vpp]),
        "orc_mod_destroy": (None, [vp]),
        "orc_mod_taps": (i, [vp, _f64p, i, ip]),
        "orc_mod_modulate_bits": (i, [vp, cp, i64, i, _f32p, i64, i64p]),
        "orc_mod_modulate_bytes": (i, [vp, _u8p, i64, _u8p, i64, _u8p, i64, i, _f32p, i64, i64p]),
        "orc_demod_create": (i, [i, i, f, i, d, d, d, i, cp, i, i64, _vpp]),
        "orc_demod_destroy": (None, [vp]),
        "orc_demod_bits": (i, [vp, _f32p, i64, cp, i64, i64p]),
        "orc_demod_bytes": (i, [vp, _f32p, i64, _u8p, i64, _u8p, i64, _u8p, i64, i64p]),
        "orc_demod_frame_bits": (i, [vp, C.c_char_p, i64, _u8p, i64, _u8p, i64, _u8p, i64, i64p]),
        "orc_demod_constellation": (i, [vp, _f32p, i64, _f32p, i64, i64p]),
        "orc_demod_loop_state": (None, [vp, _f64p, _f64p, _f64p, _f64p, _f32p, _f32p]),
        "orc_demod_in_frame": (i, [vp]),
        "orc_nco_create": (i, [d, d, d, d, u64, u64, _vpp]),
        "orc_nco_destroy": (None, [vp]),
        "orc_nco_generate": (None, [vp, _f64p, i64]),
        "orc_noise_iq": (None, [f, i64, u64, u64, u64, _f32p]),
        "orc_channel_apply": (None, [vp, vp, i, _f32p, _f32p, i64, _f32p]),
        "orc_multipath": (None, [_f32p, i64, _f32p, ip, i, _f32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args


def _check(st: int):
    if st == 0:
        return
    if st == ERR_NULL:
        raise ArgumentNullException()
    if st == ERR_ARG:
        raise ArgumentException()
    if st == ERR_RANGE:
        raise ArgumentOutOfRangeException()
    raise RuntimeError(f"oracle status {st}")


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(_f32p)


def _bytes_arr(b) -> np.ndarray:
    return np.frombuffer(bytes(b), dtype=np.uint8).copy() if len(b) else np.zeros(0, np.uint8)


def _up(a: np.ndarray):
    return a.ctypes.data_as(_u8p)


# ------------------------------------------------------------------------------------------
def rng_u64(seed, stream, counter) -> int:
    return lib().orc_rng_u64(seed, stream, counter)


def rng_double(seed, stream, counter) -> float:
    return lib().orc_rng_double(seed, stream, counter)


def fill_uniform(seed: int, stream: int, first: int, n: int) -> np.ndarray:
    out = np.empty(n, np.float32)
    lib().orc_fill_uniform(seed, stream, first, n, _fp(out))
    return out


def fill_bytes(seed: int, stream: int, first: int, n: int) -> bytes:
    out = np.empty(n, np.uint8)
    lib().orc_fill_bytes(seed, stream, first, n, _up(out))
    return out.tobytes()


class RRCFilter:
    @staticmethod
    def generateCoefficents(spanSymbols: float, beta: float, sampleRate: int, SymbolRate: int) -> np.ndarray:
        n = C.c_int(0)
        _check(lib().orc_rrc_taps(spanSymbols, beta, sampleRate, SymbolRate, None, 0, C.byref(n)))
        out = np.empty(max(n.value, 0), np.float64)
        _check(lib().orc_rrc_taps(spanSymbols, beta, sampleRate, SymbolRate, out.ctypes.data_as(_f64p), n.value, C.byref(n)))
        return out


def real_taps_to_iq(h) -> np.ndarray:
    t = np.zeros(2 * len(h), np.float32)
    t[0::2] = np.asarray(h, np.float64).astype(np.float32)
    return t


class _Handle:
    _destroy = None

    def __init__(self):
        self._h = C.c_void_p()

    def __del__(self):
        try:
            if self._h and self._destroy:
                getattr(lib(), self._destroy)(self._h)
                self._h = None
        except Exception:
            pass


class ComplexFIRFilter(_Handle):
    _destroy = "orc_fir_destroy"

    def __init__(self, tapsInterleavedIQ):
        super().__init__()
        if tapsInterleavedIQ is None:
            raise ArgumentNullException()
        t = _f32(tapsInterleavedIQ)
        self.taps = t.copy()
        _check(lib().orc_fir_create(_fp(t), t.size, C.byref(self._h)))

    def Filter(self, iqIn, out_len=None) -> np.ndarray:
        x = _f32(iqIn)
        y = np.empty(x.size if out_len is None else out_len, np.float32)
        _check(lib().orc_fir_filter(self._h, _fp(x), _fp(y), x.size, y.size))
        return y

    def fftFilter(self, iqData) -> np.ndarray:
        if iqData is None:
            raise ArgumentNullException()
        x = _f32(iqData)
        y = np.empty(x.size if x.size % 2 == 0 else 0, np.float32)
        _check(lib().orc_fir_fft_filter(self._h, _fp(x), _fp(y), x.size))
        return y

    def reset(self):
        lib().orc_fir_reset(self._h)


def decimate(filtered_iq, D: int, skip: int = 0) -> np.ndarray:
    """Row N1 (north_star item (2), SURVEY §8d "decimate-by-D variant").  The reference's matched filter is non-decimating
    (ComplexFIRFilter.Filter, MS/Models/FIRFilter.cs:80-91, called at MS/QPSKDeModulator.cs:360), so the oracle of the
    decimating filter is that output kept at stream indices skip, skip + D, skip + 2D, ...: pass Filter()'s output."""
    y = _f32(filtered_iq).reshape(-1, 2)
    return np.ascontiguousarray(y[skip::D]).reshape(-1)


def fir_filter_f64(taps_iq, iq_in) -> np.ndarray:
    t, x = _f32(taps_iq), _f32(iq_in)
    y = np.empty(x.size, np.float64)
    lib().orc_fir_filter_f64(_fp(t), t.size // 2, _fp(x), y.ctypes.data_as(_f64p), x.size // 2)
    return y


def convolve_f64(a_iq, b) -> np.ndarray:
    a = np.ascontiguousarray(a_iq, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    out = np.empty(2 * (a.size // 2 + b.size - 1), np.float64)
    lib().orc_convolve_f64(a.ctypes.data_as(_f64p), a.size // 2, b.ctypes.data_as(_f64p), b.size, out.ctypes.data_as(_f64p))
    return out


def fir_filter_mt(taps_iq, iq_in: np.ndarray, out: np.ndarray, n_floats_each: int, threads: int):
    t = _f32(taps_iq)
    _check(lib().orc_fir_filter_mt(_fp(t), t.size, _fp(iq_in), _fp(out), n_floats_each, threads))


class FLLBandEdgeFilter(_Handle):
    _destroy = "orc_fll_destroy"

    def __init__(self, sps, rolloff, filterSize, bandwidth):
        super().__init__()
        self.filterSize = filterSize
        _check(lib().orc_fll_create(sps, rolloff, filterSize, bandwidth, C.byref(self._h)))

    def taps(self):
        lo = np.empty(2 * self.filterSize, np.float32)
        up = np.empty(2 * self.filterSize, np.float32)
        lib().orc_fll_taps(self._h, _fp(lo), _fp(up))
        return lo, up

    def Process(self, inputIQ, out_len=None) -> np.ndarray:
        x = _f32(inputIQ)
        y = np.empty(x.size if out_len is None else out_len, np.float32)
        _check(lib().orc_fll_process(self._h, _fp(x), _fp(y), x.size, y.size))
        return y

    @property
    def state(self):
        p, f = C.c_float(), C.c_float()
        lib().orc_fll_get_state(self._h, C.byref(p), C.byref(f))
        return p.value, f.value

    @state.setter
    def state(self, pf):
        lib().orc_fll_set_state(self._h, pf[0], pf[1])


class MuellerMuller(_Handle):
    _destroy = "orc_mm_destroy"

    def __init__(self, samplesPerSymbol, kp, ki):
        super().__init__()
        _check(lib().orc_mm_create(samplesPerSymbol, kp, ki, C.byref(self._h)))

    def Process(self, incomingMfSamplesIQ, cap_floats=None) -> np.ndarray:
        x = _f32(incomingMfSamplesIQ)
        cap = x.size if cap_floats is None else cap_floats
        y = np.empty(max(cap, 0), np.float32)
        n = C.c_int(0)
        _check(lib().orc_mm_process(self._h, _fp(x), x.size, _fp(y), cap, C.byref(n)))
        return y[: 2 * n.value].copy()

    @property
    def state(self):
        b, q = C.c_int(), C.c_int()
        mu, integ = C.c_double(), C.c_double()
        lib().orc_mm_get_state(self._h, C.byref(b), C.byref(mu), C.byref(integ), C.byref(q))
        return dict(baseIndex=b.value, mu=mu.value, ncoIntegral=integ.value, queued=q.value)


def mm_gains_from_bw(sym_bw: float):
    kp, ki = C.c_double(), C.c_double()
    lib().orc_mm_gains_from_bw(sym_bw, C.byref(kp), C.byref(ki))
    return kp.value, ki.value


class CostasLoopQpsk(_Handle):
    _destroy = "orc_costas_destroy"

    def __init__(self, sampleRate, loopBandwidthHz, damping=0.707):
        super().__init__()
        _check(lib().orc_costas_create(sampleRate, loopBandwidthHz, damping, C.byref(self._h)))

    def Process(self, iqIn, out_len=None) -> np.ndarray:
        x = _f32(iqIn)
        y = np.empty(x.size if out_len is None else out_len, np.float32)
        _check(lib().orc_costas_process(self._h, _fp(x), _fp(y), x.size, y.size))
        return y

    def GetState(self):
        t, f = C.c_double(), C.c_double()
        lib().orc_costas_get_state(self._h, C.byref(t), C.byref(f))
        return t.value, f.value

    def gains(self):
        a, b = C.c_double(), C.c_double()
        lib().orc_costas_gains(self._h, C.byref(a), C.byref(b))
        return a.value, b.value


class BitPacker:
    @staticmethod
    def BytesToBitString(data: bytes) -> str:
        a = _bytes_arr(data)
        buf = C.create_string_buffer(a.size * 8 + 1)
        lib().orc_bytes_to_bits(_up(a), a.size, buf)
        return buf.raw[: a.size * 8].decode("ascii")

    @staticmethod
    def BitsToBytes(bits: str, bitOffset: int) -> bytes:
        b = bits.encode("ascii")
        out = np.empty(len(b) // 8 + 1, np.uint8)
        n = lib().orc_bits_to_bytes(b, len(b), bitOffset, _up(out), out.size)
        if n < 0:
            _check(int(n))
        return out[:n].tobytes()

    @staticmethod
    def IndexOf(haystack: bytes, needle: bytes) -> int:
        h, n = _bytes_arr(haystack), _bytes_arr(needle)
        return int(lib().orc_index_of(_up(h), h.size, _up(n), n.size))


class QPSKModulator(_Handle):
    _destroy = "orc_mod_destroy"

    def __init__(self, SampleRate, SymbolRate, RrcAlpha=0.9, rrcSpan=6, differentialEncoding=True, tsc=None):
        super().__init__()
        t = None if tsc is None else tsc.encode("ascii")
        _check(lib().orc_mod_create(SampleRate, SymbolRate, RrcAlpha, rrcSpan, int(differentialEncoding), t, C.byref(self._h)))

    def getCoeef(self) -> np.ndarray:
        n = C.c_int(0)
        lib().orc_mod_taps(self._h, None, 0, C.byref(n))
        out = np.empty(n.value, np.float64)
        _check(lib().orc_mod_taps(self._h, out.ctypes.data_as(_f64p), n.value, C.byref(n)))
        return out

    def Modulate(self, data: str, pulseShaping: bool = True) -> np.ndarray:
        if data is None:
            raise ArgumentNullException()
        b = data.encode("ascii")
        n = C.c_int64(0)
        _check(lib().orc_mod_modulate_bits(self._h, b, len(b), int(pulseShaping), None, 0, C.byref(n)))
        out = np.empty(n.value, np.float32)
        _check(lib().orc_mod_modulate_bits(self._h, b, len(b), int(pulseShaping), _fp(out), out.size, C.byref(n)))
        return out

    def ModulateBytes(self, payload: bytes, startMarker: bytes, endMarker: bytes, pulseShaping: bool = True) -> np.ndarray:
        p, s, e = _bytes_arr(payload), _bytes_arr(startMarker), _bytes_arr(endMarker)
        n = C.c_int64(0)
        args = (self._h, _up(p), p.size, _up(s), s.size, _up(e), e.size, int(pulseShaping))
        _check(lib().orc_mod_modulate_bytes(*args, None, 0, C.byref(n)))
        out = np.empty(n.value, np.float32)
        _check(lib().orc_mod_modulate_bytes(*args, _fp(out), out.size, C.byref(n)))
        return out

    def ModulateTextUtf8(self, text: str, startMarker="\x02", endMarker="\x03", pulseShaping=True) -> np.ndarray:
        if text is None:
            raise ArgumentNullException()
        return self.ModulateBytes(text.encode("utf-8"), startMarker.encode("utf-8"), endMarker.encode("utf-8"), pulseShaping)


class QPSKDeModulator(_Handle):
    _destroy = "orc_demod_destroy"

    def __init__(self, SampleRate, SymbolRate, RrcAlpha=0.9, rrcSpan=6, SymbolSyncBandwith=0.0001,
                 CostasLoopBandwith=120.0, CFOLoopBandwith=float(np.float32(0.0001)), differentialEncoding=True,
                 tsc=None, use_fll=False, ring_capacity=0):
        super().__init__()
        t = None if tsc is None else tsc.encode("ascii")
        _check(lib().orc_demod_create(SampleRate, SymbolRate, RrcAlpha, rrcSpan, SymbolSyncBandwith, CostasLoopBandwith,
                                      CFOLoopBandwith, int(differentialEncoding), t, int(use_fll), ring_capacity,
                                      C.byref(self._h)))

    def DeModulate(self, SamplesIQ) -> str:
        if SamplesIQ is None:
            raise ArgumentNullException()
        x = _f32(SamplesIQ)
        cap = x.size + 16
        buf = C.create_string_buffer(cap)
        n = C.c_int64(0)
        _check(lib().orc_demod_bits(self._h, _fp(x), x.size, buf, cap, C.byref(n)))
        return buf.raw[: n.value].decode("ascii")

    def DeModulateBytes(self, samplesIQ, startMarker: bytes, endMarker: bytes, cap: int = 0) -> bytes:
        x = _f32(samplesIQ)
        s, e = _bytes_arr(startMarker), _bytes_arr(endMarker)
        out = np.empty(cap or (x.size // 8 + 64), np.uint8)          # a frame may span calls: pass cap for long ones
        n = C.c_int64(0)
        _check(lib().orc_demod_bytes(self._h, _fp(x), x.size, _up(s), s.size, _up(e), e.size, _up(out), out.size, C.byref(n)))
        return out[: n.value].tobytes()

    def FrameBits(self, bits: str, startMarker: bytes, endMarker: bytes, cap: int = 0) -> bytes:
        """DeModulateBytes from the point where it holds rxBits (:179-259)."""
        s, e = _bytes_arr(startMarker), _bytes_arr(endMarker)
        raw = bits.encode("ascii")
        out = np.empty(cap or (len(raw) // 8 + 64), np.uint8)
        n = C.c_int64(0)
        _check(lib().orc_demod_frame_bits(self._h, raw, len(raw), _up(s), s.size, _up(e), e.size, _up(out), out.size, C.byref(n)))
        return out[: n.value].tobytes()

    def DeModulateTextUtf8(self, samplesIQ, startMarker="\x02", endMarker="\x03") -> str:
        p = self.DeModulateBytes(samplesIQ, startMarker.encode("utf-8"), endMarker.encode("utf-8"))
        return p.decode("utf-8", errors="replace") if p else ""

    def deModulateConstellation(self, SamplesIQ) -> np.ndarray:
        x = _f32(SamplesIQ)
        y = np.empty(x.size, np.float32)
        n = C.c_int64(0)
        _check(lib().orc_demod_constellation(self._h, _fp(x), x.size, _fp(y), y.size, C.byref(n)))
        return y[: 2 * n.value].copy()

    def loop_state(self):
        ct, cf, mu, mi = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        fp, ff = C.c_float(), C.c_float()
        lib().orc_demod_loop_state(self._h, C.byref(ct), C.byref(cf), C.byref(mu), C.byref(mi), C.byref(fp), C.byref(ff))
        return dict(costas_theta=ct.value, costas_freq=cf.value, mm_mu=mu.value, mm_integral=mi.value,
                    fll_phase=fp.value, fll_freq=ff.value)

    @property
    def in_frame(self) -> bool:
        return bool(lib().orc_demod_in_frame(self._h))


class NCO(_Handle):
    _destroy = "orc_nco_destroy"

    def __init__(self, frequencyHz, sampleRateHz, PpmInstabillity=0.0, initialPhaseRad=0.0, seed=0, stream=0):
        super().__init__()
        _check(lib().orc_nco_create(frequencyHz, sampleRateHz, PpmInstabillity, initialPhaseRad, seed, stream, C.byref(self._h)))

    def GenerateBlock(self, count: int) -> np.ndarray:
        out = np.empty(2 * count, np.float64)
        lib().orc_nco_generate(self._h, out.ctypes.data_as(_f64p), count)
        return out


def noise_iq(dbfs: float, count: int, seed: int, stream: int, first_sample: int = 0) -> np.ndarray:
    out = np.empty(2 * count, np.float32)
    lib().orc_noise_iq(dbfs, count, seed, stream, first_sample, _fp(out))
    return out


def channel_apply(tx: NCO, rx: NCO, mode: int, x_iq, noise=None) -> np.ndarray:
    x = _f32(x_iq)
    y = np.empty_like(x)
    nz = None if noise is None else _f32(noise)
    lib().orc_channel_apply(tx._h, rx._h, mode, _fp(x), None if nz is None else _fp(nz), x.size // 2, _fp(y))
    return y


def multipath(x_iq, gains_iq, delays) -> np.ndarray:
    x, g = _f32(x_iq), _f32(gains_iq)
    dl = np.ascontiguousarray(delays, np.int32)
    y = np.empty_like(x)
    lib().orc_multipath(_fp(x), x.size // 2, _fp(g), dl.ctypes.data_as(C.POINTER(C.c_int)), dl.size, _fp(y))
    return y
