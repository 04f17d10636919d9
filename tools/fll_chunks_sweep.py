import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
import qpsk_modulator_demodulator_b200 as Q
import bench_chain
Q.set_device(0)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); s = ts.cuda_stream
r = {}
for C in (1, 8, 32, 64, 128, 255, 256, 512, 1024):
    o = bench_chain.run_chain(Q, torch, None, 1, 0, s, steps=5, warmup=3, use_fll=True, channels_per_gpu=C)
    r[C] = round(o["ms_per_step"], 4)
print(os.environ.get("QPSK_DEMOD_CHUNKS", "default"), json.dumps(r))
