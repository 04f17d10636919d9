"""Randomised parity soak of the hot path against the oracle (the FLL has its own: tools/fuzz_fll.py).  Every case
draws its design parameters, sizes, channel count and chunking at random and feeds the oracle the SAME chunks:
  fir     ComplexFIRFilter.Filter / fftFilter, real and complex taps, 1..300 taps: exact mode bit-identical (streaming),
          fast / FMA / split modes and fftFilter within 1e-5 x max|y|
  mm      MuellerMuller.Process: symbols and loop state bit-identical
  costas  CostasLoopQpsk.Process: outputs bit-identical, (theta, freq) to 1e-11 (fp64 sin/cos ulps)
  demod   QPSKDeModulator.DeModulate in exact mode, with / without FLL, TSC, differential: bit strings identical
  mod     QPSKModulator.Modulate: within 1e-5 x max|y|
  framer  the framer half of DeModulateBytes on random bit chunks (qpsk_demod_frame_bits): payloads and in-frame flags identical
usage: python tools/fuzz_path.py [cases per family] [seed] [families, comma separated]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
import qpsk_modulator_demodulator_b200 as Q

O.build()
Q.set_device(0)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
fams = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fir", "mm", "costas", "demod", "mod", "framer"]
rng = np.random.default_rng(seed)
TOL = 1e-5
bad = {}


def note(fam, info):
    bad[fam] = bad.get(fam, 0) + 1
    print("MISMATCH", fam, info)


def cuts_of(L, k):
    return sorted(set([0, L] + [2 * int(v) // 2 for v in rng.integers(0, L + 1, k)]))


def bits_eq(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def close(a, b):
    return a.shape == b.shape and (b.size == 0 or np.abs(a - b).max() <= TOL * max(np.abs(b).max(), 1e-30))


def fuzz_fir():
    for k in range(cases):
        n = int(rng.choice([1, 2, 3, 7, 8, 9, 16, 17, 21, 33, 40, 65, 129, 257, int(rng.integers(1, 300))]))
        real = rng.random() < 0.6
        taps = np.zeros(2 * n, np.float32)
        taps[0::2] = rng.standard_normal(n).astype(np.float32)
        if not real:
            taps[1::2] = rng.standard_normal(n).astype(np.float32)
        C = int(rng.choice([1, 1, 2, 3, 5]))
        L = int(rng.integers(1, 6000))
        x = rng.standard_normal((C, 2 * L)).astype(np.float32)
        mode = int(rng.choice([0, 1, 1, 2, 3, 3]))          # QPSK_FIR_FAST / EXACT / FMA / SPLIT
        g = Q.ComplexFIRFilter(taps, channels=C)
        g.set_mode(mode)
        os_ = [O.ComplexFIRFilter(taps) for _ in range(C)]
        info = dict(n=n, real=real, C=C, L=L, mode=mode)
        gots, wants = [[] for _ in range(C)], [[] for _ in range(C)]
        for a, b in zip(*(lambda c: (c[:-1], c[1:]))(cuts_of(L, int(rng.integers(0, 4))))):
            xs = np.ascontiguousarray(x[:, 2 * a:2 * b])
            got = g.Filter(xs if C > 1 else xs[0])
            got = got if C > 1 else got[None, :]
            for c in range(C):
                want = os_[c].Filter(xs[c])
                gots[c].append(got[c])
                wants[c].append(want)
                if mode == Q.FIR_EXACT and not bits_eq(got[c], want):
                    note("fir", dict(info, a=a, b=b, c=c))
        if mode != Q.FIR_EXACT:                             # tolerance on the stream's own scale, not on a short chunk's
            for c in range(C):
                if not close(np.concatenate(gots[c]), np.concatenate(wants[c])):
                    note("fir", dict(info, c=c))
        # stateless form on a fresh block
        xs = rng.standard_normal((C, 2 * int(rng.integers(1, 3000)))).astype(np.float32)
        got = g.fftFilter(xs if C > 1 else xs[0])
        got = got if C > 1 else got[None, :]
        for c in range(C):
            if not close(got[c], os_[c].fftFilter(xs[c])):
                note("fir.fft", dict(info, c=c))


def fuzz_mm():
    for k in range(cases):
        sps = float(rng.choice([2.0, 2.0, 4.0, 8.0, 2.5, 3.7, 16.0]))
        kp, ki = O.mm_gains_from_bw(float(10 ** rng.uniform(-5, -1.5)))
        C = int(rng.choice([1, 1, 2, 33]))
        L = int(rng.integers(1, 5000))
        # a plausible matched-filter output: noisy +-1 corners held for sps samples with a timing drift
        t = (np.arange(L) / (sps * (1 + rng.uniform(-0.01, 0.01)))).astype(np.int64)
        x = np.zeros((C, 2 * L), np.float32)
        for c in range(C):
            sym = rng.choice([-0.7, 0.7], (t.max() + 1, 2))
            x[c, 0::2] = sym[t, 0] + 0.1 * rng.standard_normal(L)
            x[c, 1::2] = sym[t, 1] + 0.1 * rng.standard_normal(L)
        g = Q.MuellerMuller(sps, kp, ki, channels=C)
        os_ = [O.MuellerMuller(sps, kp, ki) for _ in range(C)]
        info = dict(sps=sps, kp=kp, C=C, L=L)
        cuts = cuts_of(L, int(rng.integers(0, 4)))
        for a, b in zip(cuts[:-1], cuts[1:]):
            xs = np.ascontiguousarray(x[:, 2 * a:2 * b])
            got = g.Process(xs if C > 1 else xs[0])
            got = got if C > 1 else [got]
            for c in range(C):
                if not bits_eq(got[c], os_[c].Process(xs[c])):
                    note("mm", dict(info, a=a, b=b, c=c))
        st = g.state
        for c in range(C):
            w = os_[c].state
            gotst = {kk: (v if C == 1 else v[c]) for kk, v in st.items()}
            if any(float(gotst[kk]) != float(w[kk]) for kk in ("baseIndex", "mu", "ncoIntegral", "queued")):
                note("mm.state", dict(info, c=c, got=gotst, want=w))


def fuzz_costas():
    for k in range(cases):
        fs = float(rng.choice([1000.0, 5e6, 48000.0]))
        bw = fs / float(rng.choice([60.0, 120.0, 1000.0]))        # (a loop wider than ~fs/30 is chaotic: fp64 sincos ulps diverge)
        damping = float(rng.choice([0.707, 0.5, 1.0]))
        C = int(rng.choice([1, 1, 2, 40]))
        L = int(rng.integers(1, 4000))
        x = np.zeros((C, 2 * L), np.float32)
        for c in range(C):
            ph = rng.uniform(-3, 3) + rng.uniform(-0.02, 0.02) * np.arange(L)
            s = (rng.choice([-1, 1], L) + 1j * rng.choice([-1, 1], L)) * np.exp(1j * ph) * rng.uniform(0.05, 3)
            s = s + 0.05 * (rng.standard_normal(L) + 1j * rng.standard_normal(L))
            x[c, 0::2], x[c, 1::2] = s.real, s.imag
        g = Q.CostasLoopQpsk(fs, bw, damping, channels=C)
        os_ = [O.CostasLoopQpsk(fs, bw, damping) for _ in range(C)]
        info = dict(fs=fs, bw=bw, damping=damping, C=C, L=L)
        cuts = cuts_of(L, int(rng.integers(0, 4)))
        for a, b in zip(cuts[:-1], cuts[1:]):
            xs = np.ascontiguousarray(x[:, 2 * a:2 * b])
            got = g.Process(xs if C > 1 else xs[0])
            got = got if C > 1 else got[None, :]
            for c in range(C):
                if not bits_eq(got[c], os_[c].Process(xs[c])):
                    note("costas", dict(info, a=a, b=b, c=c))
        th, fr = g.GetState()
        for c in range(C):
            w = os_[c].GetState()
            gs = (float(th), float(fr)) if C == 1 else (float(th[c]), float(fr[c]))
            # fp64 state: the device's sin/cos is within 1 ulp of glibc's, visible at ~1e-16 relative
            if not np.allclose(gs, (float(w[0]), float(w[1])), rtol=1e-11, atol=1e-13):
                note("costas.state", dict(info, c=c, got=gs, want=w))


TSC = "1011000111010010" * 4


def fuzz_demod():
    for k in range(cases):
        sps = int(rng.choice([2, 2, 4, 8]))
        rs = int(rng.choice([1000, 250000]))
        fs = rs * sps
        alpha = float(np.float32(rng.choice([0.2, 0.35, 0.4, 0.9])))
        span = int(rng.choice([4, 6, 10]))
        diff = bool(rng.random() < 0.7)
        tsc = None if rng.random() < 0.4 else TSC[: int(rng.choice([16, 64]))]
        use_fll = bool(rng.random() < 0.4)
        C = int(rng.choice([1, 1, 2, 5, 33, 40]))       # >= 32 channels: the matched filter runs fused inside the symbol-stage kernel
        mod = O.QPSKModulator(fs, rs, alpha, span, diff, tsc)
        rows = []
        nb = 2 * int(rng.integers(50, 1500))
        reps = int(rng.integers(1, 3))
        for c in range(C):
            bits = "".join(rng.choice(["0", "1"], nb))
            y = np.concatenate([mod.Modulate(bits, True) for _ in range(reps)])
            n = y.size // 2
            ph = rng.uniform(-3, 3) + rng.uniform(-2e-3, 2e-3) * np.arange(n)
            z = (y[0::2] + 1j * y[1::2]) * np.exp(1j * ph) + 0.02 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
            r = np.empty(2 * n, np.float32)
            r[0::2], r[1::2] = z.real, z.imag
            rows.append(r)
        x = np.stack(rows)
        L = x.shape[1] // 2
        kw = dict(RrcAlpha=alpha, rrcSpan=span, SymbolSyncBandwith=float(rng.choice([1e-4, 2e-3])), differentialEncoding=diff, tsc=tsc,
                  use_fll=use_fll)
        gd = Q.QPSKDeModulator(fs, rs, channels=C, **kw)
        gd.set_fir_mode(Q.FIR_EXACT)
        ods = [O.QPSKDeModulator(fs, rs, **kw) for _ in range(C)]
        info = dict(sps=sps, alpha=alpha, span=span, diff=diff, tsc=None if tsc is None else len(tsc), fll=use_fll, C=C, L=L)
        cuts = cuts_of(L, int(rng.integers(0, 3)))
        for a, b in zip(cuts[:-1], cuts[1:]):
            xs = np.ascontiguousarray(x[:, 2 * a:2 * b])
            got = gd.DeModulate(xs if C > 1 else xs[0])
            got = got if C > 1 else [got]
            for c in range(C):
                if got[c] != ods[c].DeModulate(xs[c]):
                    note("demod", dict(info, a=a, b=b, c=c))


def fuzz_mod():
    for k in range(cases):
        sps = int(rng.choice([2, 3, 4, 8, 16]))
        rs = 1000
        alpha = float(np.float32(rng.uniform(0.05, 1.0)))
        span = int(rng.integers(2, 17))
        diff = bool(rng.random() < 0.6)
        tsc = None if rng.random() < 0.5 else TSC[: int(rng.integers(1, 65))]
        nb = int(rng.integers(0, 9000))
        bits = "".join(rng.choice(["0", "1"], nb))
        shaping = bool(rng.random() < 0.85)
        gm, om = Q.QPSKModulator(rs * sps, rs, alpha, span, diff, tsc), O.QPSKModulator(rs * sps, rs, alpha, span, diff, tsc)
        got, want = gm.Modulate(bits, shaping), om.Modulate(bits, shaping)
        ok = close(got, want) if shaping else bits_eq(got, want)
        if not ok:
            note("mod", dict(sps=sps, alpha=alpha, span=span, diff=diff, tsc=None if tsc is None else len(tsc), nb=nb, shaping=shaping))


def fuzz_framer():
    bits_of = lambda b: "".join(format(v, "08b") for v in b)
    for k in range(cases):
        sm = rng.integers(0, 256, int(rng.choice([1, 1, 2, 5, 13, 40])), dtype=np.uint8).tobytes()
        em = rng.integers(0, 256, int(rng.choice([1, 2, 4, 9])), dtype=np.uint8).tobytes()
        C = int(rng.choice([1, 1, 3, 7, 33, 70]))
        ring = int(rng.choice([8, 32, 1 << 20]))
        rows = []
        for c in range(C):
            t = ""
            for _f in range(int(rng.integers(1, 5))):
                t += "".join(rng.choice(["0", "1"], int(rng.integers(0, 60))))
                body = rng.integers(0, 256, int(rng.integers(0, 50)), dtype=np.uint8).tobytes()
                if rng.random() < 0.2:
                    body = body[: len(body) // 2] + sm + body[len(body) // 2:]
                frame = bits_of(sm + body + em)
                if rng.random() < 0.15:
                    frame = frame[: int(rng.integers(0, len(frame)))]          # a frame cut short
                t += frame
            rows.append(t + "".join(rng.choice(["0", "1"], int(rng.integers(0, 60)))))
        gd = Q.QPSKDeModulator(4000, 1000, max_frame_bytes=ring, channels=C)
        ods = [O.QPSKDeModulator(4000, 1000, ring_capacity=ring) for _ in range(C)]
        pos = [0] * C
        info = dict(ns=len(sm), ne=len(em), C=C, ring=ring)
        while any(pos[c] < len(rows[c]) for c in range(C)):
            hi = int(rng.choice([20, 90, 600]))
            take = [int(rng.integers(0, hi)) for _ in range(C)]
            chunks = [rows[c][pos[c]:pos[c] + take[c]] for c in range(C)]
            pos = [pos[c] + take[c] for c in range(C)]
            got = gd.FrameBits(chunks if C > 1 else chunks[0], sm, em, cap=4096)
            got = got if C > 1 else [got]
            inf = np.atleast_1d(gd.in_frame)
            for c in range(C):
                if got[c] != ods[c].FrameBits(chunks[c], sm, em, cap=4096) or bool(inf[c]) != bool(ods[c].in_frame):
                    note("framer", dict(info, c=c, pos=pos[c]))


for f in fams:
    {"framer": fuzz_framer, "fir": fuzz_fir, "mm": fuzz_mm, "costas": fuzz_costas, "demod": fuzz_demod, "mod": fuzz_mod}[f]()
    print(f"fuzz_path[{f}]: {cases} cases, {bad.get(f, 0) + sum(v for k2, v in bad.items() if k2.startswith(f + '.'))} mismatching", flush=True)
sys.exit(1 if bad else 0)
