"""Chain end-to-end leg (host samples in, payload bytes out) against the number of time chunks of the copy pipeline.
usage: python tools/e2e_chunks_sweep.py   (spawns one process per setting: the knob is read once per process)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    import qpsk_modulator_demodulator_b200 as Q
    import bench_chain
    Q.set_device(0)
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
    r = bench_chain.run_chain_e2e(Q, torch, None, 1, 0, ts.cuda_stream, steps=7, use_fll=(sys.argv[2] == "fll"))
    print(json.dumps({k: (round(r[k]["value"]), r[k]["ms_per_step"]) for k in ("pinned", "registered", "pageable", "cs16_pinned")}))
else:
    for fll in ("nofll", "fll"):
        for n in (0, 1, 2, 3, 4, 6, 8, 16):
            env = dict(os.environ)
            if n:
                env["QPSK_DEMOD_HOST_CHUNKS"] = str(n)
            out = subprocess.run([sys.executable, __file__, "child", fll], env=env, capture_output=True, text=True)
            print(fll, "chunks", n or "default", out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
