"""FIR host call on pageable arrays after different things have run in the process (which of them slows the host copy pool?).
usage: python tools/pageable_probe.py [none|torchcpu|oracle|affinity]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qpsk_modulator_demodulator_b200 as Q
Q.set_device(0)
what = sys.argv[1] if len(sys.argv) > 1 else "none"
n = 1 << 26
if what == "torchcpu":
    a = torch.randn(1 << 22)
    b = torch.cat([a, a]).abs().max().item()
if what == "oracle":
    import oracle as O
    t = O.real_taps_to_iq(O.RRCFilter.generateCoefficents(16, 0.35, 16000, 1000))
    O.ComplexFIRFilter(t).Filter(np.zeros(2 * 8192, np.float32))
print(what, "affinity before:", len(os.sched_getaffinity(0)), "threads:", len(os.listdir("/proc/self/task")))
taps = Q.real_taps_to_iq(Q.RRCFilter.generateCoefficents(16, 0.35, 2000, 1000))
f = Q.ComplexFIRFilter(taps)
x = np.random.default_rng(1).standard_normal(2 * n).astype(np.float32)
y = np.empty_like(x); y.fill(0)
f.Filter(x[: 1 << 22], y[: 1 << 22]); f.reset()
t0 = time.perf_counter(); f.Filter(x, y); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(what, "pageable FIR", round(n / dt / 1e6), "Msamples/s", "threads now:", len(os.listdir("/proc/self/task")))
