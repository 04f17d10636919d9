"""Throughput of the FMA FIR kernel with complex taps (generic ComplexFIRFilter, band-edge filter shapes)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qpsk_modulator_demodulator_b200 as Q
Q.set_device(0)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); s = ts.cuda_stream
L = 1 << 27
x = torch.rand(2 * L, device="cuda") - 0.5
y = torch.empty_like(x)
out = {}
rng = np.random.default_rng(1)
peak = float(os.environ.get("FMA_PEAK_TFLOPS", "70.7"))
for n in (33, 40, 65, 129, 257):
    taps = (rng.standard_normal(2 * n) / np.sqrt(n)).astype(np.float32)
    f = Q.ComplexFIRFilter(taps)
    for _ in range(2):
        f.filter_dev(x.data_ptr(), y.data_ptr(), 2 * L, stream=s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    for _ in range(3):
        f.filter_dev(x.data_ptr(), y.data_ptr(), 2 * L, stream=s)
    e1.record(ts); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    tf = 8.0 * n * L / (ms * 1e-3) / 1e12
    out[str(n)] = {"ms": ms, "gsamples_s": L / ms / 1e6, "tflops": tf, "frac_of_fma_peak": tf / peak,
                   "hbm_gbs": 16.0 * L / (ms * 1e-3) / 1e9}
print(json.dumps(out))
