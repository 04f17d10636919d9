"""Throughput of the FIR in QPSK_FIR_EXACT mode (reference summation order, no FMA) against the default FAST mode."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qpsk_modulator_demodulator_b200 as Q
Q.set_device(0)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); s = ts.cuda_stream
L = 1 << 26
x = torch.rand(2 * L, device="cuda") - 0.5
y = torch.empty_like(x)
out = {}
for span, sps in ((16, 2), (16, 4), (16, 16)):
    h = Q.RRCFilter.generateCoefficents(span, 0.35, sps * 1000, 1000)
    taps = Q.real_taps_to_iq(h)
    for mode, name in ((Q.FIR_FAST, "fast"), (Q.FIR_EXACT, "exact")):
        f = Q.ComplexFIRFilter(taps)
        f.set_mode(mode)
        for _ in range(2):
            f.filter_dev(x.data_ptr(), y.data_ptr(), 2 * L, stream=s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        for _ in range(3):
            f.filter_dev(x.data_ptr(), y.data_ptr(), 2 * L, stream=s)
        e1.record(ts); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out[f"{taps.size // 2}_{name}"] = {"ms": ms, "gsamples_s": L / ms / 1e6}
print(json.dumps(out))
