"""One launch of the decimating FIR per (taps, D) configuration of bench.py's `decimate` leg — for an ncu capture.
usage: python tools/decim_probe.py [log2_samples]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qpsk_modulator_demodulator_b200 as Q
Q.set_device(0)
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); s = ts.cuda_stream
x = torch.empty(2 * n, dtype=torch.float32, device="cuda")
y = torch.empty(2 * n, dtype=torch.float32, device="cuda")
Q.fill_uniform_dev(1, 0, 0, 2 * n, x.data_ptr(), s)
out = []
for span, sps, dec in ((16, 2, 2), (16, 4, 2), (16, 8, 2), (16, 4, 4), (16, 8, 8), (16, 16, 16)):
    t = Q.real_taps_to_iq(Q.RRCFilter.generateCoefficents(span, 0.35, sps * 1000, 1000))
    f = Q.ComplexFIRFilter(t)
    f.decimate_dev(x.data_ptr(), 2 * n, dec, y.data_ptr(), 2 * n, stream=s)
    torch.cuda.synchronize()
    out.append((t.size // 2, dec, f.last_kernel()))
print(json.dumps(out))
