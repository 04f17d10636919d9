"""numpy twin of the CPU oracle — TEST INFRASTRUCTURE (parity unpinned, see qpsk_oracle.cpp header).

A second, independently written restatement of the same reference code, read from the C# sources
(not from qpsk_oracle.cpp) and kept deliberately naive: numpy.float32 scalars/arrays for the fp32
paths (every operation rounds separately), Python floats + libm for the fp64 paths, a real FFT for
fftFilter.  tests/test_oracle_twin.py requires the C++ oracle and this twin to agree, which guards
against transcription mistakes in either.  Pure-Python loops: small cases only.

"MS/" = /root/reference/Modulation-Simulation/.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
LANES = 8                      # Vector<float>.Count on AVX2
PI_F = f32(math.pi)            # MathF.PI
TWO_PI_F = f32(2.0) * PI_F     # MS/Models/Band-Edge Filter.cs:16


def cs_round(x: float) -> int:
    """Math.Round(double): round half to even."""
    return int(round(x))       # Python's round() is banker's rounding too


def cr_sinf(x) -> np.float32:
    """MathF.Sin model: correctly rounded fp32 (fp64 evaluation, one rounding)."""
    return f32(math.sin(float(x)))


def cr_cosf(x) -> np.float32:
    return f32(math.cos(float(x)))


# ---- MS/Models/RRC-filter.cs:16-75 -----------------------------------------------------------
def rrc_taps(span_symbols: float, beta: float, sample_rate: int, symbol_rate: int) -> np.ndarray:
    sps = cs_round(sample_rate / symbol_rate)
    taps = cs_round(span_symbols) * sps + 1
    h = [0.0] * max(taps, 0)
    mid = (taps - 1) // 2 if taps > 0 else 0
    pi, eps = math.pi, 1e-8
    for n in range(taps):
        t = (n - mid) / float(sps)
        if abs(t) < eps:
            val = 1.0 + beta * (4.0 / pi - 1.0)
        elif abs(abs(t) - 1.0 / (4.0 * beta)) < eps:
            val = (beta / math.sqrt(2.0)) * ((1.0 + 2.0 / pi) * math.sin(pi / (4.0 * beta)) +
                                             (1.0 - 2.0 / pi) * math.cos(pi / (4.0 * beta)))
        else:
            num = math.sin(pi * t * (1.0 - beta)) + 4.0 * beta * t * math.cos(pi * t * (1.0 + beta))
            fbt = 4.0 * beta * t
            den = pi * t * (1.0 - fbt * fbt)     # Math.Pow(x, 2.0) modelled as the correctly rounded square
            val = num / den
        h[n] = val
    energy = 0.0
    for v in h:
        energy += v * v
    norm = math.sqrt(energy)
    return np.array([v / norm for v in h], np.float64)


def real_taps_to_iq(h) -> np.ndarray:
    t = np.zeros(2 * len(h), np.float32)
    t[0::2] = np.asarray(h, np.float64).astype(np.float32)
    return t


# ---- MS/Models/FIRFilter.cs:8-232 -------------------------------------------------------------
class ComplexFIRFilter:
    def __init__(self, taps_iq):
        t = np.asarray(taps_iq, np.float32)
        if t.size % 2 or t.size == 0:
            raise ValueError("taps")
        self.taps = t.copy()
        self.n = t.size // 2
        self.hI = t[0::2][::-1].copy()          # tapsRev[k] = taps[N-1-k]  (:43-48)
        self.hQ = t[1::2][::-1].copy()
        self.dI = np.zeros(2 * self.n, np.float32)
        self.dQ = np.zeros(2 * self.n, np.float32)
        self.pos = 0

    def _dot(self, start):                       # ComplexDotWindow :144-211, hardware-accelerated branch
        N = self.n
        xI, xQ = self.dI[start:start + N], self.dQ[start:start + N]
        n_vec = N - (N % LANES)
        vI = np.zeros(LANES, np.float32)
        vQ = np.zeros(LANES, np.float32)
        for i in range(0, n_vec, LANES):
            hi, hq = self.hI[i:i + LANES], self.hQ[i:i + LANES]
            xi, xq = xI[i:i + LANES], xQ[i:i + LANES]
            vI = vI + ((hi * xi) - (hq * xq))
            vQ = vQ + ((hi * xq) + (hq * xi))
        aI, aQ = f32(0), f32(0)
        for lane in range(LANES):
            aI = aI + vI[lane]
            aQ = aQ + vQ[lane]
        for i in range(n_vec, N):
            aI = aI + ((self.hI[i] * xI[i]) - (self.hQ[i] * xQ[i]))
            aQ = aQ + ((self.hI[i] * xQ[i]) + (self.hQ[i] * xI[i]))
        return aI, aQ

    def filter1(self, inI, inQ):                 # :59-77
        p, N = self.pos, self.n
        self.dI[p] = inI; self.dQ[p] = inQ
        self.dI[p + N] = inI; self.dQ[p + N] = inQ
        start = p + 1
        if start >= N:
            start -= N
        out = self._dot(start)
        p += 1
        self.pos = 0 if p == N else p
        return out

    def Filter(self, iq):                        # :80-91
        x = np.asarray(iq, np.float32)
        y = np.empty_like(x)
        for s in range(0, x.size, 2):
            y[s], y[s + 1] = self.filter1(x[s], x[s + 1])
        return y

    def fftFilter(self, iq):                     # :96-141, through an actual fp64 FFT like MathNet
        x = np.asarray(iq, np.float32)
        n_data = x.size // 2
        if n_data == 0:
            return np.zeros(0, np.float32)
        n_conv = n_data + self.n - 1
        size = 1
        while size < n_conv:
            size <<= 1
        X = np.fft.fft(x[0::2].astype(np.float64) + 1j * x[1::2].astype(np.float64), size)
        H = np.fft.fft(self.taps[0::2].astype(np.float64) + 1j * self.taps[1::2].astype(np.float64), size)
        c = np.fft.ifft(X * H)[self.n - 1: self.n - 1 + n_data]
        y = np.empty(2 * n_data, np.float32)
        y[0::2] = c.real.astype(np.float32)
        y[1::2] = c.imag.astype(np.float32)
        return y


# ---- MS/Models/Band-Edge Filter.cs:14-203 -----------------------------------------------------
def _sinc(x: np.float32) -> np.float32:
    if x == f32(0):
        return f32(1)
    arg = PI_F * x
    return cr_sinf(arg) / arg


class FLLBandEdgeFilter:
    def __init__(self, sps, rolloff, size, bandwidth):
        self.sps, self.rolloff, self.size, self.bw = f32(sps), f32(rolloff), int(size), f32(bandwidth)
        self.phase, self.freq = f32(0), f32(0)
        self.alpha = f32(0)
        self.beta = f32(4.0) * self.bw / self.sps
        self.max_freq = TWO_PI_F * (f32(2.0) / self.sps)
        self.min_freq = -self.max_freq
        mid = (self.size - 1) // 2
        bb = np.zeros(self.size, np.float32)
        total = f32(0)
        for i in range(self.size):
            k = f32(i - mid) / (f32(2.0) * self.sps)
            pos = self.rolloff * k
            tap = _sinc(pos - f32(0.5)) + _sinc(pos + f32(0.5))
            total = total + tap
            bb[i] = tap
        for i in range(self.size):
            bb[i] = bb[i] / total
        lo = np.zeros(2 * self.size, np.float32)
        up = np.zeros(2 * self.size, np.float32)
        for i in range(self.size):
            k = f32(i - mid) / (f32(2.0) * self.sps)
            angle = -TWO_PI_F * (f32(1.0) + self.rolloff) * k
            li, lq = bb[i] * cr_cosf(angle), bb[i] * cr_sinf(angle)
            lo[2 * i], lo[2 * i + 1] = li, lq
            up[2 * i], up[2 * i + 1] = li, -lq
        self.lower_taps, self.upper_taps = lo, up
        self.lower, self.upper = ComplexFIRFilter(lo), ComplexFIRFilter(up)

    def process1(self, inI, inQ):
        c, s = cr_cosf(self.phase), cr_sinf(self.phase)
        outI = inI * c - inQ * s
        outQ = inI * s + inQ * c
        upI, upQ = self.upper.filter1(outI, outQ)
        loI, loQ = self.lower.filter1(outI, outQ)
        pow_up = upI * upI + upQ * upQ
        pow_lo = loI * loI + loQ * loQ
        err = pow_lo - pow_up
        self.freq = self.freq + self.beta * err
        self.phase = self.phase + (self.freq + self.alpha * err)
        if self.phase > TWO_PI_F or self.phase < -TWO_PI_F:
            self.phase = f32(math.remainder(float(self.phase), float(TWO_PI_F)))   # exact in fp64, then exact in fp32
        if self.freq > self.max_freq:
            self.freq = self.max_freq
        elif self.freq < self.min_freq:
            self.freq = self.min_freq
        return outI, outQ

    def Process(self, iq):
        x = np.asarray(iq, np.float32)
        y = np.empty_like(x)
        for s in range(0, x.size, 2):
            y[s], y[s + 1] = self.process1(x[s], x[s + 1])
        return y


# ---- MS/Models/MuellerMuller.cs:17-250 --------------------------------------------------------
class MuellerMuller:
    def __init__(self, sps, kp, ki):
        self.sps, self.kp, self.ki = float(sps), float(kp), float(ki)
        self.base, self.mu, self.integ = 1, 0.0, 0.0
        self.pS = (f32(0), f32(0))
        self.pD = (f32(0), f32(0))
        self.has_prev = False
        self.buf = np.zeros(0, np.float32)

    def _interp(self, n, mu):
        b = self.buf
        xm1, x0, x1, x2 = (b[2 * (n - 1):2 * n], b[2 * n:2 * n + 2], b[2 * n + 2:2 * n + 4], b[2 * n + 4:2 * n + 6])
        t = f32(mu)
        tm1, tm2, tp1 = t - f32(1), t - f32(2), t + f32(1)
        sixth, half = f32(1) / f32(6), f32(1) / f32(2)
        c_m1 = -(t * tm1 * tm2) * sixth
        c_0 = (tp1 * tm1 * tm2) * half
        c_1 = -(tp1 * t * tm2) * half
        c_2 = (tp1 * t * tm1) * sixth
        oI = c_m1 * xm1[0] + c_0 * x0[0] + c_1 * x1[0] + c_2 * x2[0]
        oQ = c_m1 * xm1[1] + c_0 * x0[1] + c_1 * x1[1] + c_2 * x2[1]
        return f32(oI), f32(oQ)

    def Process(self, iq, cap_floats=None):
        x = np.asarray(iq, np.float32)
        cap = x.size if cap_floats is None else cap_floats
        self.buf = np.concatenate([self.buf, x])
        count = self.buf.size // 2
        out = []
        while self.base + 2 < count:
            cI, cQ = self._interp(self.base, self.mu)
            dI = f32(1) if cI >= 0 else f32(-1)
            dQ = f32(1) if cQ >= 0 else f32(-1)
            if self.has_prev:
                t1 = float(self.pD[0]) * float(cI) + float(self.pD[1]) * float(cQ)
                t2 = float(dI) * float(self.pS[0]) + float(dQ) * float(self.pS[1])
                e = t1 - t2
                self.integ += self.ki * e
                corr = self.kp * e + self.integ
                corr = min(corr, 0.1)
                corr = max(corr, -0.1)
                adv = self.sps + corr
            else:
                self.has_prev = True
                adv = self.sps
            o = len(out)
            if o + 1 >= cap:
                break
            out += [cI, cQ]
            self.pS, self.pD = (cI, cQ), (dI, dQ)
            new_time = self.base + self.mu + adv
            self.base = int(math.floor(new_time))
            self.mu = new_time - self.base
            if self.base + 1 >= count:
                break
        consumed = min(max(0, self.base - 1), max(0, count - 3))
        if consumed > 0:
            self.buf = self.buf[2 * consumed:]
            self.base -= consumed
        return np.array(out, np.float32)


# ---- MS/Models/CostasLoopQpsk.cs:19-131 -------------------------------------------------------
class CostasLoopQpsk:
    def __init__(self, fs, bw_hz, damping=0.707):
        bw = 2.0 * math.pi * bw_hz / fs
        d = 1.0 + 2.0 * damping * bw + bw * bw
        self.alpha, self.beta = (4.0 * damping * bw) / d, (4.0 * bw * bw) / d
        self.theta = self.freq = 0.0

    def process1(self, inI, inQ):
        c, s = math.cos(self.theta), math.sin(self.theta)
        mi = float(inI) * c + float(inQ) * s
        mq = float(inQ) * c - float(inI) * s
        oI, oQ = f32(mi), f32(mq)
        eI = 1.0 if oI >= 0 else -1.0
        eQ = 1.0 if oQ >= 0 else -1.0
        pe = eI * mq - eQ * mi
        self.freq += self.beta * pe
        self.theta += self.freq + self.alpha * pe
        if self.theta > math.pi:
            self.theta -= 2.0 * math.pi
        elif self.theta < -math.pi:
            self.theta += 2.0 * math.pi
        return oI, oQ

    def Process(self, iq):
        x = np.asarray(iq, np.float32)
        y = np.empty_like(x)
        for s in range(0, x.size, 2):
            y[s], y[s + 1] = self.process1(x[s], x[s + 1])
        return y


# ---- MS/Models/HelperFunctions.cs:11-71 -------------------------------------------------------
def bytes_to_bit_string(data: bytes) -> str:
    return "".join(f"{b:08b}" for b in data)


def bits_to_bytes(bits: str, bit_offset: int) -> bytes:
    usable = len(bits) - bit_offset
    if usable < 8:
        return b""
    n = usable // 8
    return bytes(int(bits[bit_offset + 8 * i: bit_offset + 8 * i + 8].replace(" ", "0"), 2) if set(bits) <= {"0", "1"}
                 else sum((1 if bits[bit_offset + 8 * i + j] == "1" else 0) << (7 - j) for j in range(8)) for i in range(n))


# ---- MS/QPSKModulator.cs:18-168 ---------------------------------------------------------------
INV_SQRT2 = f32(0.7071067811865475)


class QPSKModulator:
    def __init__(self, fs, rs, alpha=0.9, span=6, diff=True, tsc=None):
        self.fs, self.rs, self.diff = fs, rs, diff
        self.tsc = None if (tsc is None or tsc.strip() == "") else tsc
        self.coeff = rrc_taps(span, alpha, fs, rs)
        self.rrc = ComplexFIRFilter(real_taps_to_iq(self.coeff))

    def Modulate(self, data: str, pulse=True) -> np.ndarray:
        if self.tsc is not None:
            data = self.tsc + data
        nd = len(data) >> 1
        if nd == 0:
            return np.zeros(0, np.float32)
        sps = self.fs // self.rs
        delay = (len(self.coeff) - 1) // 2
        base = delay + nd * sps
        total = base + delay if pulse else base
        up = np.zeros(2 * total, np.float32)
        pI, pQ = INV_SQRT2, INV_SQRT2
        w = delay
        for d in range(nd):
            bi, bq = ord(data[2 * d]) - 48, ord(data[2 * d + 1]) - 48
            if self.diff:
                if bi == 0 and bq == 0:
                    dI, dQ = f32(1), f32(0)
                elif bi == 0 and bq == 1:
                    dI, dQ = f32(0), f32(1)
                elif bi == 1 and bq == 1:
                    dI, dQ = f32(-1), f32(0)
                else:
                    dI, dQ = f32(0), f32(-1)
                sI = pI * dI - pQ * dQ
                sQ = pI * dQ + pQ * dI
                pI, pQ = sI, sQ
            else:
                sI = -INV_SQRT2 if bi == 0 else INV_SQRT2
                sQ = -INV_SQRT2 if bq == 0 else INV_SQRT2
            up[2 * w], up[2 * w + 1] = sI, sQ
            w += sps
        return self.rrc.fftFilter(up) if pulse else up

    def ModulateBytes(self, payload: bytes, sm: bytes, em: bytes, pulse=True):
        if len(sm) == 0 or len(em) == 0:
            raise ValueError("marker")
        return self.Modulate(bytes_to_bit_string(sm + payload + em), pulse)


# ---- MS/QPSKDeModulator.cs:11-456 -------------------------------------------------------------
class QPSKDeModulator:
    def __init__(self, fs, rs, alpha=0.9, span=6, sym_bw=0.0001, costas_bw=120.0, cfo_bw=float(np.float32(0.0001)),
                 diff=True, tsc=None, use_fll=False):
        self.diff, self.use_fll = diff, use_fll
        self.tsc = None if (tsc is None or tsc.strip() == "") else tsc
        self.rrc = ComplexFIRFilter(real_taps_to_iq(rrc_taps(span, float(np.float32(alpha)), fs, rs)))
        self.fll = FLLBandEdgeFilter(float(fs // rs), alpha, 40, float(np.float32(cfo_bw)))
        zeta = 1.0 / math.sqrt(2.0)
        wn = ((2.0 * math.pi * sym_bw) / (zeta + 0.25) / zeta)
        den = 1.0 + 2.0 * zeta * wn + wn * wn
        self.mm = MuellerMuller(fs / float(rs), (4.0 * zeta * wn) / den, (4.0 * wn * wn) / den)
        self.costas = CostasLoopQpsk(float(rs), rs / costas_bw)
        self.have_prev, self.prev = False, (f32(0), f32(0))
        self.in_frame, self.carry, self.ring = False, "", bytearray()
        self.pack_byte, self.pack_bits = 0, 0

    def DeModulate(self, iq) -> str:
        x = np.asarray(iq, np.float32)
        if x.size == 0:
            return ""
        if self.use_fll:
            x = self.fll.Process(x)
        sym = self.mm.Process(self.rrc.Filter(x), x.size)
        bits = []
        for k in range(sym.size // 2):
            rI, rQ = self.costas.process1(sym[2 * k], sym[2 * k + 1])
            dI = f32(1) if rI >= 0 else f32(-1)
            dQ = f32(1) if rQ >= 0 else f32(-1)
            if self.diff:
                if not self.have_prev:
                    self.prev, self.have_prev = (dI, dQ), True
                    continue
                deI = dI * self.prev[0] + dQ * self.prev[1]
                deQ = dQ * self.prev[0] - dI * self.prev[1]
                self.prev = (dI, dQ)
                if abs(deI) >= abs(deQ):
                    bits.append("00" if deI >= 0 else "11")
                else:
                    bits.append("01" if deQ >= 0 else "10")
            else:
                if dI < 0:
                    bits.append("00" if dQ < 0 else "01")
                else:
                    bits.append("11" if dQ >= 0 else "10")
        rx = "".join(bits)
        if self.tsc is not None:
            idx = rx.find(self.tsc)
            return "" if idx < 0 else rx[idx + len(self.tsc):]
        return rx

    def _reset(self):
        self.in_frame, self.carry, self.ring = False, "", bytearray()
        self.pack_byte, self.pack_bits = 0, 0

    def _append(self, bits: str) -> int:
        produced = 0
        for c in bits:
            self.pack_byte = ((self.pack_byte << 1) | (1 if c == "1" else 0)) & 0xFF
            self.pack_bits += 1
            if self.pack_bits == 8:
                self.ring.append(self.pack_byte)
                produced += 1
                self.pack_bits, self.pack_byte = 0, 0
        return produced

    def DeModulateBytes(self, iq, sm: bytes, em: bytes) -> bytes:
        if len(sm) == 0 or len(em) == 0:
            raise ValueError("marker")
        return self.FrameBits(self.DeModulate(iq), sm, em)

    def FrameBits(self, rx: str, sm: bytes, em: bytes) -> bytes:
        """DeModulateBytes from the point where it holds rxBits (MS/QPSKDeModulator.cs:179-259)."""
        if len(sm) == 0 or len(em) == 0:
            raise ValueError("marker")
        if rx == "":
            return b""
        if not self.in_frame:
            cand = self.carry + rx
            for off in range(8):
                by = bits_to_bytes(cand, off)
                if len(by) == 0:
                    continue
                s = by.find(sm)
                if s < 0:
                    continue
                end_bit = off + 8 * (s + len(sm))
                if end_bit > len(cand):
                    continue
                self.in_frame = True
                self.ring, self.pack_byte, self.pack_bits = bytearray(), 0, 0
                appended = self._append(cand[end_bit:])
                at = bytes(self.ring).find(em, max(0, len(self.ring) - (appended + len(em))))
                if at >= 0:
                    out = bytes(self.ring[:at])
                    self._reset()
                    return out
                return b""
            keep = min(len(cand), len(sm) * 8 + 7)
            self.carry = "" if keep == 0 else cand[len(cand) - keep:]
            return b""
        appended = self._append(rx)
        at = bytes(self.ring).find(em, max(0, len(self.ring) - (appended + len(em))))
        if at >= 0:
            out = bytes(self.ring[:at])
            self._reset()
            return out
        return b""


def save_as_cs16(iq) -> tuple:
    """HelperFunctions.SaveAsCs16 (MS/Models/HelperFunctions.cs:75-106) without the file write: interleaved fp32 IQ
    (the Complex[] the reference takes is built from such floats, TB/HelperModels.cs:49-58) -> (int16 I,Q,..., maxVal).
    maxVal = max(|re|, |im|) (:83-90), 1.0 if below 1e-12 (:92); each component = (short) clamp(v / maxVal *
    short.MaxValue) in fp64 (:97-103) — the C# cast truncates toward zero."""
    x = np.asarray(iq, np.float32)
    if x.size == 0:
        raise ValueError("IQ array is empty.")                     # :79-80
    maxv = 0.0
    for v in x.tolist():                                           # python floats = fp64
        a = abs(v)
        if a > maxv:
            maxv = a
    norm = maxv if maxv >= 1e-12 else 1.0
    out = np.empty(x.size, np.int16)
    for k, v in enumerate(x.tolist()):
        s = v / norm * 32767.0
        s = max(-32768.0, min(32767.0, s))
        out[k] = int(s)                                            # int() truncates toward zero
    return out, maxv
