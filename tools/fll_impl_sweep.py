"""FLL kernel choice by stream count: demodulator chain with the FLL on (QPSK_FLL_IMPL / QPSK_FLL_PAIRS from the environment),
ms per step at 2048 ... 16384 channels.  usage: [QPSK_FLL_IMPL=group] python tools/fll_impl_sweep.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qpsk_modulator_demodulator_b200 as Q
import bench_chain
Q.set_device(0)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); s = ts.cuda_stream
r = {}
for C in [int(v) for v in os.environ.get("SWEEP", "2048,4096,8192,16384").split(",")]:
    o = bench_chain.run_chain(Q, torch, None, 1, 0, s, steps=3, warmup=2, use_fll=True, channels_per_gpu=C, parity_channels=0)
    r[C] = round(o["ms_per_step"], 3)
print(os.environ.get("QPSK_FLL_IMPL", "duo"), os.environ.get("QPSK_FLL_PAIRS", "auto"), json.dumps(r))
