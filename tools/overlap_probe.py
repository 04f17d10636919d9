"""Does a pinned host->device copy overlap the FLL kernel?  FLL alone, copy alone (contiguous / 2-D row pieces through
cudaMemcpy2DAsync as the library issues them), and the chunk pipeline copy(t+1) || FLL(t) rebuilt here from the same calls."""
import ctypes, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qpsk_modulator_demodulator_b200 as Q
Q.set_device(0)
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
C, L = 2048, 4380
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
x = torch.randn((C, 2 * L), dtype=torch.float32, device="cuda") * 0.3
y = torch.empty_like(x)
z = torch.empty_like(x)
fll = Q.FLLBandEdgeFilter(2.0, 0.4, 40, 0.01, channels=C)
pin = Q.PinnedBuffer(C * 2 * L)
hx = torch.from_numpy(pin.array).reshape(C, 2 * L)
row = 2 * L * 4
def cuts(chunks):
    step = -(-L // chunks)
    return [(n0, min(step, L - n0)) for n0 in range(0, L, step)]
def copy2d(n0, ln, stream):
    r = rt.cudaMemcpy2DAsync(z.data_ptr() + 8 * n0, row, pin.array.ctypes.data + 8 * n0, row, 8 * ln, C, 1, ctypes.c_void_p(stream.cuda_stream))
    assert r == 0, r
def run_fll(chunks=1):
    for n0, ln in cuts(chunks):
        fll.process_dev(x.data_ptr() + 8 * n0, y.data_ptr() + 8 * n0, 2 * ln, 2 * L, 2 * L, sb.cuda_stream)
def run_copy(chunks=1):
    for n0, ln in cuts(chunks):
        copy2d(n0, ln, sa)
def pipeline(chunks):
    evs = []
    for n0, ln in cuts(chunks):
        copy2d(n0, ln, sa)
        e = torch.cuda.Event(); e.record(sa); evs.append(e)
    for (n0, ln), e in zip(cuts(chunks), evs):
        sb.wait_event(e)
        fll.process_dev(z.data_ptr() + 8 * n0, y.data_ptr() + 8 * n0, 2 * ln, 2 * L, 2 * L, sb.cuda_stream)
def t(fn, n=9):
    for _ in range(2): fn()
    torch.cuda.synchronize(); v = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); v.append(1e3 * (time.perf_counter() - t0))
    return round(sorted(v)[n // 2], 3)
res = {"fll x1": t(run_fll), "fll x6": t(lambda: run_fll(6)), "copy contiguous (torch)": t(lambda: z.copy_(hx, non_blocking=True))}
for k in (1, 2, 4, 6, 16):
    res[f"copy2d x{k}"] = t(lambda: run_copy(k))
for k in (2, 4, 6):
    res[f"copy || fll, independent, x{k}"] = t(lambda: (run_copy(k), run_fll(k)))
    res[f"pipeline x{k}"] = t(lambda: pipeline(k))
print(json.dumps(res, indent=0))
