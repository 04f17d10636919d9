"""Different handles on different host threads at the same time — the reference's usage pattern: the TX and RX threads of
TB/SDR/ModDemodOverSDR.cs:94-183 each own their modulator / demodulator objects.  ctypes releases the GIL inside the C
calls, so the threads really overlap in the library; every result must still equal the oracle's."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TSC = "11001010011101100100100110101100" + "01110100111001011010001101101001"
FS = 10_000_000
RS = FS // 2
ALPHA = float(np.float32(0.4))


def test_concurrent_handles_match_oracle(gpu, orc):
    n_threads, rounds = 6, 12
    rng = np.random.default_rng(77)
    jobs = []
    for t in range(n_threads):
        taps = np.zeros(2 * (17 + 8 * t), np.float32)
        taps[0::2] = rng.standard_normal(taps.size // 2).astype(np.float32)
        xs = [rng.standard_normal(2 * int(rng.integers(100, 3000))).astype(np.float32) for _ in range(rounds)]
        payloads = [rng.integers(0, 256, int(rng.integers(20, 200)), dtype=np.uint8).tobytes() for _ in range(rounds)]
        jobs.append((taps, xs, payloads))
    # oracle results first (single thread)
    want = []
    for taps, xs, payloads in jobs:
        of = orc.ComplexFIRFilter(taps)
        om = orc.QPSKModulator(FS, RS, ALPHA, 10, True, TSC)
        od = orc.QPSKDeModulator(FS, RS, ALPHA, 10, tsc=TSC)
        w = []
        for x, p in zip(xs, payloads):
            y = of.Filter(x)
            burst = om.ModulateBytes(p, b"S", b"E")
            w.append((y, burst, od.DeModulate(burst)))
        want.append(w)
    errors = []
    start = threading.Barrier(n_threads)

    def worker(t):
        try:
            taps, xs, payloads = jobs[t]
            gf = gpu.ComplexFIRFilter(taps)
            gf.set_mode(gpu.FIR_EXACT)
            gm = gpu.QPSKModulator(FS, RS, ALPHA, 10, True, TSC)
            gd = gpu.QPSKDeModulator(FS, RS, ALPHA, 10, tsc=TSC)
            gd.set_fir_mode(gpu.FIR_EXACT)
            start.wait()
            for k, (x, p) in enumerate(zip(xs, payloads)):
                y = gf.Filter(x)
                burst = gm.ModulateBytes(p, b"S", b"E")
                wy, wb, wbits = want[t][k]
                if not np.array_equal(y.view(np.uint32), wy.view(np.uint32)):
                    errors.append((t, k, "fir"))
                if burst.shape != wb.shape or np.abs(burst - wb).max() > 1e-5 * np.abs(wb).max():
                    errors.append((t, k, "mod"))
                if gd.DeModulate(wb) != wbits:                       # the oracle's burst: identical inputs on both sides
                    errors.append((t, k, "demod"))
        except Exception as e:                                       # noqa: BLE001 - reported below
            errors.append((t, repr(e)))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errors, errors[:5]


def test_concurrent_pageable_streams_share_the_host_copy_pool(gpu, orc):
    """Long streams in PAGEABLE numpy arrays from several threads at once: every call goes through its handle's page-locked
    staging slots, filled and drained by the one host copy pool of the process (core.cu HostCopyPool) — the calls serialise on
    the pool, never mix their chunks, and give the one-shot result of the same kernel."""
    n_threads = 3
    L = 2 * (1 << 21) + 4567                                      # three pipeline chunks
    rng = np.random.default_rng(5)
    taps = [orc.real_taps_to_iq(orc.RRCFilter.generateCoefficents(10, 0.35, (2 + t) * 1000, 1000)) for t in range(n_threads)]
    xs = [rng.standard_normal(2 * L).astype(np.float32) for _ in range(n_threads)]
    got = [None] * n_threads
    errors = []
    start = threading.Barrier(n_threads)

    def worker(t):
        try:
            f = gpu.ComplexFIRFilter(taps[t])
            f.set_mode(gpu.FIR_FMA)
            start.wait()
            got[t] = f.Filter(xs[t])
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errors, errors
    import torch
    for t in range(n_threads):
        dx = torch.from_numpy(xs[t]).cuda()
        dy = torch.empty_like(dx)
        f = gpu.ComplexFIRFilter(taps[t])
        f.set_mode(gpu.FIR_FMA)
        f.filter_dev(dx.data_ptr(), dy.data_ptr(), 2 * L)         # one launch over the whole stream, device resident
        torch.cuda.synchronize()
        assert np.array_equal(dy.cpu().numpy().view(np.uint32), got[t].view(np.uint32)), t
        # and a window against the oracle
        w = orc.ComplexFIRFilter(taps[t]).Filter(xs[t][: 2 * 50000])
        assert np.abs(got[t][: 2 * 50000] - w).max() <= 1e-5 * np.abs(w).max()
