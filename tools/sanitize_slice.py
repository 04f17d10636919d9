"""A compact pass over every hand-rolled pipeline of libqpskcuda.so, meant to run under compute-sanitizer
(memcheck / racecheck / synccheck / initcheck) — SURVEY §5's counterpart of the reference's (absent) race detection:

  FIR      fir_tma_kernel (mbarrier ring, TMA bulk loads / per-warp bulk stores) over several tile seams, streaming in
           two chunks (delay-line ping-pong), real and complex taps, stateless alignment; fir_exact_real_kernel;
           fir_split2_kernel; fir_dec2_kernel / fir_decim_tma_kernel (D = 2, 4, 8, 16)
  FLL      fll_duo_kernel (chain warp / side warp hand-off through shared memory + mbarriers), 40 and 10 taps,
           fll_lane_kernel forced, generic group kernel (13 taps)
  symsync  symsync_decode_kernel (cp.async double buffer, two-warp symbol queue), with TSC strip and framer
  chain    the time-chunk pipeline (two streams, events) with the FLL on
  mod      mod_tile_rot / mod_shape kernels; channel simulator; BER counters; pack / unpack; CS16

Every result is also compared with the oracle, so a run that the sanitizer slows down is still a parity run.
usage: compute-sanitizer --tool <tool> python tools/sanitize_slice.py [small]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle as O
import qpsk_modulator_demodulator_b200 as Q

small = len(sys.argv) > 1 and sys.argv[1] == "small"        # racecheck is ~100x: fewer samples, same kernels
Q.set_device(0)
TSC = "11001010011101100100100110101100" + "01110100111001011010001101101001"
done = []


def close(a, b, tol=1e-5):
    return a.shape == b.shape and float(np.abs(a - b).max()) <= tol * max(float(np.abs(b).max()), 1e-30)


# ---- FIR ---------------------------------------------------------------------------------------------------------
L = 6000 if small else 30000                                  # tile = 2560 samples: 3 / 12 tiles, ragged last tile
x = O.fill_uniform(3, 0, 0, 2 * L)
for span, sps in ((16, 2), (16, 4)) if small else ((16, 2), (16, 4), (16, 16)):
    taps = Q.real_taps_to_iq(Q.RRCFilter.generateCoefficents(span, 0.35, sps * 1000, 1000))
    want = O.ComplexFIRFilter(taps).Filter(x)
    f = Q.ComplexFIRFilter(taps)
    cut = 2 * (L // 3 + 1)
    got = np.concatenate([f.Filter(x[:cut]), f.Filter(x[cut:])])
    assert close(got, want), ("fir fast", span, sps)
    fe = Q.ComplexFIRFilter(taps)
    fe.set_mode(Q.FIR_EXACT)
    ge = np.concatenate([fe.Filter(x[:cut]), fe.Filter(x[cut:])])
    assert np.array_equal(ge.view(np.uint32), want.view(np.uint32)), ("fir exact", span, sps)
    assert close(Q.ComplexFIRFilter(taps).fftFilter(x), O.ComplexFIRFilter(taps).fftFilter(x)), ("fftFilter", span, sps)
# the 2-parallel split kernel (S planes per warp, single-buffered output slice) and the decimators on the TMA ring
t129 = Q.real_taps_to_iq(Q.RRCFilter.generateCoefficents(16, 0.35, 8000, 1000))
fs_ = Q.ComplexFIRFilter(t129)
fs_.set_mode(Q.FIR_SPLIT)
cut = 2 * (L // 3 + 1)
gs = np.concatenate([fs_.Filter(x[:cut]), fs_.Filter(x[cut:])])
assert "split2" in fs_.last_kernel() and close(gs, O.ComplexFIRFilter(t129).Filter(x)), "fir split"
for dec in (2, 4, 8, 16):
    fd = Q.ComplexFIRFilter(t129)
    gd_ = np.concatenate([fd.Decimate(x[:cut], dec), fd.Decimate(x[cut:], dec)])
    assert close(gd_, O.decimate(O.ComplexFIRFilter(t129).Filter(x), dec)), ("decimate", dec, fd.last_kernel())
rng = np.random.default_rng(5)
ct = rng.standard_normal(2 * 40).astype(np.float32) * 0.1
assert close(Q.ComplexFIRFilter(ct).Filter(x), O.ComplexFIRFilter(ct).Filter(x)), "fir complex taps"
xb = np.ascontiguousarray(x[: 2 * 3 * 1000].reshape(3, 2000))
gb = Q.ComplexFIRFilter(taps, channels=3).Filter(xb)
for c in range(3):
    assert close(gb[c], O.ComplexFIRFilter(taps).Filter(xb[c])), "fir batch"
done.append("fir")

# ---- FLL ---------------------------------------------------------------------------------------------------------
Lf = 600 if small else 3000
for size, C in ((40, 5), (10, 3), (13, 2)):
    xf = (0.5 * rng.standard_normal((C, 2 * Lf))).astype(np.float32)
    g = Q.FLLBandEdgeFilter(2.0, 0.4, size, 0.01, channels=C)
    got = np.concatenate([g.Process(np.ascontiguousarray(xf[:, :400])), g.Process(np.ascontiguousarray(xf[:, 400:]))], axis=1)
    for c in range(C):
        o = O.FLLBandEdgeFilter(2.0, 0.4, size, 0.01)
        assert np.array_equal(got[c].view(np.uint32), o.Process(xf[c]).view(np.uint32)), ("fll", size, c)
done.append("fll duo/group")

# ---- modulator -> channel -> demodulator (symsync fused, TSC strip, framer), with and without the FLL ---------------
fs, rs, alpha = 10_000_000, 5_000_000, float(np.float32(0.4))
C = 40 if small else 70                                        # > 32 and > 64: partial CTAs of the 32-channel kernels
npay = 48 if small else 160
pays = [O.fill_bytes(11, 4 * c + 3, 0, npay) for c in range(C)]
om = O.QPSKModulator(fs, rs, alpha, 10, True, TSC)
gm = Q.QPSKModulator(fs, rs, alpha, 10, True, TSC)
tx = np.stack([om.ModulateBytes(p, b"S", b"E") for p in pays])
assert close(gm.ModulateBytes(pays[0], b"S", b"E"), tx[0]), "modulator"
rxs = []
for c in range(C):
    a, b = O.NCO(100e6, fs, 1, seed=21, stream=4 * c), O.NCO(100e6, fs, 1, seed=21, stream=4 * c + 1)
    rxs.append(O.channel_apply(a, b, 0, tx[c]))
rx = np.stack(rxs)
for use_fll in (False, True):
    gd = Q.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC, use_fll=use_fll, channels=C)
    ods = [O.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC, use_fll=use_fll) for _ in range(C)]
    for rep in range(3):                                       # state carried burst to burst
        got = gd.DeModulateBytes(rx, b"S", b"E")
        for c in range(C):
            assert got[c] == ods[c].DeModulateBytes(rx[c], b"S", b"E"), ("demod bytes", use_fll, rep, c)
    gb = Q.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC, use_fll=use_fll, channels=C)
    ob = [O.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC, use_fll=use_fll) for _ in range(C)]
    bits = gb.DeModulate(rx)
    for c in range(C):
        assert bits[c] == ob[c].DeModulate(rx[c]), ("demod bits", use_fll, c)
done.append("modulator / demodulator chain (fused symsync, TSC strip, framer, time-chunk pipeline)")

# ---- long single stream through the FLL time-chunk pipeline (two streams + events) ---------------------------------
Ls = 2 * (2100 if small else 9000)
xs = np.ascontiguousarray(np.tile(rx[1], 8)[:Ls])
g1 = Q.QPSKDeModulator(fs, rs, alpha, 10, tsc=None, use_fll=True)
o1 = O.QPSKDeModulator(fs, rs, alpha, 10, tsc=None, use_fll=True)
assert g1.DeModulate(xs) == o1.DeModulate(xs), "time-chunk pipeline"
done.append("fll time-chunk pipeline")

# ---- device-side helpers: channel simulator, BER counters, packing, CS16 ------------------------------------------------
import torch

ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
s = ts.cuda_stream
Cc = 9
ff = tx.shape[1]
dtx = torch.from_numpy(np.ascontiguousarray(tx[:Cc])).cuda()
dy = torch.empty((Cc, ff), dtype=torch.float32, device="cuda")
ch = Q.SimChannel(100e6, 100e6, fs, 1, 1, noise_dbfs=-40.0, mode=1, path_gains_iq=(1.0, 0.0, 0.12, 0.08), path_delays=(0, 3), seed=2026,
                  channels=Cc, first_channel=0)
ch.apply_dev(dtx.data_ptr(), ff, ff, dy.data_ptr(), ff, s)
torch.cuda.synchronize()
assert np.isfinite(dy.cpu().numpy()).all()
v = np.ascontiguousarray(rx[0])
cs, mx = Q.SaveAsCs16(v) if hasattr(Q, "SaveAsCs16") else (None, None)
done.append("channel simulator, CS16")

print("sanitize_slice ok:", "; ".join(done), "| launches:", Q.launch_count())
