import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qpsk_modulator_demodulator_b200 as Q
import bench_chain
Q.set_device(0)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); s = ts.cuda_stream
which = sys.argv[1]
ST = int(sys.argv[2]) if len(sys.argv) > 2 else 2   # timed steps
if which == 'stream':
    r = bench_chain.run_stream(Q)
elif which == 'sweep':
    r = {}
    for C in (1024, 2048, 4096, 8192, 16384):
        for fll in (False, True):
            o = bench_chain.run_chain(Q, torch, None, 1, 0, s, steps=3, warmup=2, use_fll=fll, channels_per_gpu=C, parity_channels=0)
            r[f"{C}{'_fll' if fll else ''}"] = {"msamples_s": o["value"], "ms": o["ms_per_step"], "launches": o["gpu_launches"]}
elif which == 'chain':
    r = bench_chain.run_chain(Q, torch, None, 1, 0, s, steps=ST, warmup=3, use_fll=False)
elif which == 'fll':
    r = bench_chain.run_chain(Q, torch, None, 1, 0, s, steps=ST, warmup=3, use_fll=True)
else:
    r = bench_chain.run_modulator(Q, torch, None, 1, 0, s, steps=2, warmup=2)
print(json.dumps(r))
