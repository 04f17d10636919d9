"""Launch-level look at one radio stream: pushes a few blocks through StreamingDemodulator so that
`ncu --metrics gpu__time_duration.sum` lists the kernels one block costs; without ncu prints wall time per block.
usage: python tools/stream_probe.py [blocks] [use_fll]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
import qpsk_modulator_demodulator_b200 as Q
from bench_chain import TSC

O.build()
Q.set_device(0)
blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 200
use_fll = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
fs = 10_000_000
rs = fs // 2
alpha = float(np.float32(0.4))
S, E = b"S", b"E"
mod = O.QPSKModulator(fs, rs, alpha, 10, True, TSC)
tx, rx = O.NCO(100e6, fs, 1, seed=9, stream=0), O.NCO(100e6, fs, 1, seed=9, stream=1)
bursts = [O.channel_apply(tx, rx, 0, mod.ModulateBytes(O.fill_bytes(2026, 3, 1000 * k, 512), S, E)) for k in range(8)]
n = bursts[0].size
d = Q.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC, use_fll=use_fll)
st = Q.StreamingDemodulator(d, S, E, max_block_floats=n, max_payload_bytes=2048, depth=4)
for i in range(8):
    st.push(bursts[i % 8])
st.drain()
t0 = time.perf_counter()
ok = 0
tp = 0.0
for i in range(blocks):
    t1 = time.perf_counter()
    st.push(bursts[i % 8])
    tp += time.perf_counter() - t1
    while True:
        p = st.poll()
        if p is None:
            break
        ok += bool(p)
ok += sum(bool(p) for p in st.drain())
dt = time.perf_counter() - t0
print(f"stream_probe: fll={use_fll} {blocks} blocks x {n // 2} samples: {dt / blocks * 1e6:.1f} us/block "
      f"({blocks * (n // 2) / dt / 1e6:.2f} Msamples/s), host time inside push {tp / blocks * 1e6:.1f} us/block, frames {ok}")
