"""Times the impairment simulator (qpsk_chan_apply_dev) on device-resident bursts.  usage: python tools/chan_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qpsk_modulator_demodulator_b200 as Q

Q.set_device(0)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); s = ts.cuda_stream
for C, paths in ((2048, False), (2048, True), (16384, True)):
    L = 4196
    x = torch.randn((C, 2 * L), dtype=torch.float32, device="cuda")
    y = torch.empty_like(x)
    kw = dict(path_gains_iq=(1.0, 0.0, 0.12, 0.08), path_delays=(0, 3)) if paths else {}
    ch = Q.SimChannel(100e6, 100e6, 10_000_000, 1, 1, noise_dbfs=-40.0, mode=1, seed=1, channels=C, **kw)
    for _ in range(2):
        ch.apply_dev(x.data_ptr(), 2 * L, 2 * L, y.data_ptr(), 2 * L, s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ch.apply_dev(x.data_ptr(), 2 * L, 2 * L, y.data_ptr(), 2 * L, s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"chan_apply: {C} channels x {L} samples, multipath={paths}: {ms:.3f} ms ({C * L / ms / 1e6:.2f} Gsample/s)")
