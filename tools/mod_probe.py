"""The modulator leg of bench.py at a chosen frame count, alone: CUDA-event timing, or a short run for an ncu capture.
usage: python tools/mod_probe.py [frames] [steps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qpsk_modulator_demodulator_b200 as Q
import bench_chain
Q.set_device(0)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
r = bench_chain.run_modulator(Q, torch, None, 1, 0, ts.cuda_stream, steps=steps, warmup=3, frames_per_gpu=frames)
print(json.dumps({k: r[k] for k in ("value", "ms_per_step", "roofline")}))
