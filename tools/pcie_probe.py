#!/usr/bin/env python
"""Host<->device copy-rate probe: NUMA topology of the box, the GPU's node, and pinned-copy bandwidth
(H2D alone, D2H alone, both at once) with the pinned buffers placed on each NUMA node in turn.
Explains the e2e number of bench.py (which is bound by these copies, not by the kernels)."""
import ctypes
import glob
import json
import os
import subprocess
import sys
import time

import torch

libc = ctypes.CDLL(None, use_errno=True)
SYS_set_mempolicy = 238  # x86_64
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 0, 1, 2


def set_mempolicy(mode, node=None):
    if node is None:
        return libc.syscall(SYS_set_mempolicy, MPOL_DEFAULT, None, 0)
    mask = ctypes.c_ulong(1 << node)
    return libc.syscall(SYS_set_mempolicy, mode, ctypes.byref(mask), 64)


def bw(nbytes, fn, reps=3):
    best = 0.0
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = max(best, nbytes / (time.perf_counter() - t0) / 1e9)
    return best


def main():
    out = {}
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    out["numa_nodes"] = nodes
    out["cpus"] = os.cpu_count()
    try:
        out["affinity"] = sorted(os.sched_getaffinity(0))
    except Exception:
        pass
    bus = torch.cuda.get_device_properties(0)
    try:
        pci = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id,pcie.link.gen.current,pcie.link.width.current,"
                              "pcie.link.gen.max", "--format=csv,noheader", "-i", "0"], capture_output=True, text=True).stdout.strip()
        out["pci"] = pci
        bus_id = pci.split(",")[0].strip().lower()
        # nvidia-smi prints an 8-digit domain; sysfs uses 4
        if len(bus_id.split(":")[0]) == 8:
            bus_id = bus_id[4:]
        p = f"/sys/bus/pci/devices/{bus_id}/numa_node"
        out["gpu_numa_node"] = int(open(p).read()) if os.path.exists(p) else None
        out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[-1500:]
    except Exception as e:
        out["pci_err"] = repr(e)
    n = 1 << 30
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for node in [None] + nodes:
        rc = set_mempolicy(MPOL_BIND, node)
        h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_in.fill_(1); h_out.fill_(2)
        set_mempolicy(MPOL_DEFAULT, None)

        def h2d():
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)

        def d2h():
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)

        def both():
            h2d(); d2h()

        res[str(node)] = {"rc": rc, "h2d_gbs": bw(n, h2d), "d2h_gbs": bw(n, d2h), "duplex_each_gbs": bw(n, both)}
        del h_in, h_out
    out["pinned_copy"] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
