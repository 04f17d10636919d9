"""Where a host-source DeModulateBytes call with the FLL spends its time: the device-resident call (demod_bytes_dev) on the
same bursts, the host call, and the bare copy.  usage: python tools/e2e_breakdown.py [fll|nofll]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qpsk_modulator_demodulator_b200 as Q
from bench_chain import TSC
Q.set_device(0)
use_fll = not (len(sys.argv) > 1 and sys.argv[1] == "nofll")
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); s = ts.cuda_stream
SM, EM = b"MESSAGE_START", b"MESSAGE_STOP"
fs, rs, alpha, C, npay = 10_000_000, 5_000_000, float(np.float32(0.4)), 2048, 512
mod = Q.QPSKModulator(fs, rs, alpha, 10, True, TSC)
pay = torch.empty((C, npay), dtype=torch.uint8, device="cuda")
Q.fill_bytes_dev(2026, 0, C, npay, pay.data_ptr(), s)
ff = mod.frame_floats(npay, SM, EM)
tx = torch.empty((C, ff), dtype=torch.float32, device="cuda")
mod.modulate_frames_dev(pay.data_ptr(), npay, C, SM, EM, tx.data_ptr(), ff, s)
chan = Q.SimChannel(100e6, 100e6, fs, 1, 1, noise_dbfs=-40.0, mode=1, path_gains_iq=(1.0, 0.0, 0.12, 0.08), path_delays=(0, 3), seed=2026, channels=C, first_channel=0)
rx = torch.empty((C, ff), dtype=torch.float32, device="cuda")
chan.apply_dev(tx.data_ptr(), ff, ff, rx.data_ptr(), ff, s)
torch.cuda.synchronize()
cap = 2 * (npay + 64)
dem = Q.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC, use_fll=use_fll, channels=C, max_frame_bytes=cap)
d_out = torch.empty((C, cap), dtype=torch.uint8, device="cuda")
d_nb = torch.empty(C, dtype=torch.int64, device="cuda")
res = {}
def timeit(name, fn, n=7):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    v = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); v.append(1e3 * (time.perf_counter() - t0))
    res[name] = round(sorted(v)[len(v) // 2], 3)
timeit("device bytes_dev", lambda: dem.demod_bytes_dev(rx.data_ptr(), ff, ff, SM, EM, d_out.data_ptr(), cap, d_nb.data_ptr(), s))
d_bits = torch.empty((C, dem.bits_bound(ff) + 16), dtype=torch.uint8, device="cuda")
timeit("device bits_dev", lambda: dem.demod_bits_dev(rx.data_ptr(), ff, ff, d_bits.data_ptr(), d_bits.shape[1], d_nb.data_ptr(), s))
pin = Q.PinnedBuffer(C * ff)
torch.from_numpy(pin.array).copy_(rx.reshape(-1)); torch.cuda.synchronize()
hx = torch.from_numpy(pin.array)
timeit("bare H2D (one contiguous copy)", lambda: rx.reshape(-1).copy_(hx, non_blocking=True))
out = np.zeros((C, cap), np.uint8); nb = np.zeros(C, np.int64)
dem2 = Q.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC, use_fll=use_fll, channels=C, max_frame_bytes=cap)
timeit("host demod_bytes", lambda: dem2.demod_bytes_host_ptr(pin.array.ctypes.data, ff, SM, EM, out, nb, None))
res["samples"] = C * ff // 2
print(json.dumps(res))
