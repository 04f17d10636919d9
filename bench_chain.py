"""Secondary benchmark legs of bench.py (reported under "chain" and "modulator" in the JSON line).

chain      BASELINE.json configs[2]/[3]: `channels_per_gpu` independent impaired QPSK bursts per GPU
           (2048 per GPU -> 16384 on 8 GPUs) through the batched demodulator chain
           [FLL ->] RRC MF -> Mueller-Muller -> Costas -> differential decode -> TSC strip, then the
           per-channel BER counters, gathered over all ranks with one all_gather (NCCL).
           Parameters: testAtDataLevel.cs (fs 10 MHz, Rs 5 MHz, alpha .4, span 10, 64-bit TSC,
           two 100 MHz / 1 ppm LOs) with 512-byte random payloads and -40 dBFS AWGN.
modulator  BASELINE.json configs[4]: 2048 frames x 64 KiB per GPU (1 GiB on 8 GPUs) of random payload,
           START|payload|END + TSC, differential, sps 4, span 10, alpha .35 -> polyphase RRC shaping.

Both are timed with CUDA events on the launching stream, inputs resident in HBM; the channel
simulator and payload generator run outside the timed region.  K distinct burst sets are cycled so
that every timed step reads inputs that are not L2-resident.
"""
from __future__ import annotations

import numpy as np

TSC = "11001010011101100100100110101100" + "01110100111001011010001101101001"


def _oracle_replay(sets_host, order, channels, fs, rs, alpha, use_fll):
    """The oracle's restatement of QPSKDeModulator.DeModulate on all host cores, fed — per checked channel — exactly the
    sequence of bursts the GPU demodulator saw (warm-up and timed steps, state carried from burst to burst).  Returns the
    bit string of the LAST burst per channel, the wall time and the samples processed: it is both the parity checker of
    the benchmarked run and the CPU baseline of this leg (`kind: "port"`)."""
    import os
    import time
    from concurrent.futures import ThreadPoolExecutor
    import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    last = [None] * len(channels)

    def work(t):
        done = 0
        for j in range(t, len(channels), cores):
            c = channels[j]
            d = O.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC, use_fll=use_fll)
            b = ""
            for k in order:
                b = d.DeModulate(sets_host[k][c])
                done += sets_host[k][c].size // 2
            last[j] = b
        return done

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:            # the oracle library releases the GIL
        samples = sum(ex.map(work, range(cores)))
    dt = time.perf_counter() - t0
    return last, dt, samples, cores


def _count_errors(bits: str, ref: np.ndarray) -> int:
    """qpsk_ber_count_dev's rule on the host: Hamming distance over min(n_rx, n_ref) + the bits that never arrived."""
    rx = np.frombuffer(bits.encode("ascii"), np.uint8) & 1
    n = min(rx.size, ref.size)
    return int((rx[:n] != ref[:n]).sum()) + max(0, ref.size - rx.size)


def run_chain(Q, torch, dist, world, rank, stream, steps=3, warmup=3, channels_per_gpu=2048, use_fll=False,
              n_payload=512, hbm_peak=6461.8, cpu=False, comm=None, parity_channels=None, fir_mode=None, label=None):
    """One chain leg.  `fir_mode` None = the demodulator's default matched-filter arithmetic (QPSK_FIR_EXACT).
    parity_channels: how many of this rank's channels are replayed through the oracle and compared bit for bit with what
    the TIMED run left in its output buffers (None = all when `cpu`, else 0)."""
    from qpsk_modulator_demodulator_b200 import shard
    fs = 10_000_000
    rs = fs // 2
    alpha = float(np.float32(0.4))
    C = channels_per_gpu
    total_channels = C * world
    first, last = shard.channel_range(rank, world, total_channels)
    assert last - first == C
    seed = 2026
    mod = Q.QPSKModulator(fs, rs, alpha, 10, True, TSC)
    pay = torch.empty((C, n_payload), dtype=torch.uint8, device="cuda")
    Q.fill_bytes_dev(seed, first, C, n_payload, pay.data_ptr(), stream)
    ff = mod.frame_floats(n_payload, b"S", b"E")
    tx = torch.empty((C, ff), dtype=torch.float32, device="cuda")
    mod.modulate_frames_dev(pay.data_ptr(), n_payload, C, b"S", b"E", tx.data_ptr(), ff, stream)
    # BASELINE configs[3]: unstable LOs (1 ppm each at 100 MHz: CFO + drift), AWGN, static multipath (a weak echo 3 samples late)
    chan = Q.SimChannel(100e6, 100e6, fs, 1, 1, noise_dbfs=-40.0, mode=1, path_gains_iq=(1.0, 0.0, 0.12, 0.08), path_delays=(0, 3),
                        seed=seed, channels=C, first_channel=first)
    # enough distinct burst sets that consecutive steps never hit L2 (126 MB)
    set_bytes = C * ff * 4
    K = max(2, min(8, int(np.ceil(300e6 / set_bytes))))
    rx = [torch.empty((C, ff), dtype=torch.float32, device="cuda") for _ in range(K)]
    for k in range(K):
        chan.apply_dev(tx.data_ptr(), ff, ff, rx[k].data_ptr(), ff, stream)
    nref = 8 * (n_payload + 2)
    framed = torch.cat([torch.full((C, 1), ord("S"), dtype=torch.uint8, device="cuda"), pay,
                        torch.full((C, 1), ord("E"), dtype=torch.uint8, device="cuda")], dim=1).contiguous()
    ref = torch.empty((C, nref), dtype=torch.uint8, device="cuda")
    Q.unpack_bits_dev(framed.data_ptr(), n_payload + 2, n_payload + 2, C, ref.data_ptr(), nref, stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(mode):
        """warm-up + timed steps on a fresh demodulator; returns (ms, launches, bits, n_bits, counters) of the last step"""
        dem = Q.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC, use_fll=use_fll, channels=C)
        if mode is not None:
            dem.set_fir_mode(mode)
        cap = dem.bits_bound(ff)
        bits = torch.zeros((C, cap), dtype=torch.uint8, device="cuda")
        nb = torch.zeros(C, dtype=torch.int64, device="cuda")
        cnt = torch.zeros((C, 2), dtype=torch.int32, device="cuda")

        def step(i):
            r = rx[i % K]
            dem.demod_bits_dev(r.data_ptr(), ff, ff, bits.data_ptr(), cap, nb.data_ptr(), stream)
            Q.ber_count_dev(bits.data_ptr(), cap, nb.data_ptr(), ref.data_ptr(), nref, nref, C, cnt.data_ptr(), stream)

        for i in range(warmup):
            step(i)
        barrier()
        Q.launch_count_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(warmup + i)
        e1.record()
        barrier()
        launches = Q.launch_count()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), int(launches), bits, nb, cnt

    ms, launches, bits, nb, cnt = timed(fir_mode)
    # the one collective of the path: per-channel BER counters, through the library's own NCCL communicator when there is
    # one (qpsk_ber_gather, what a C# host calls), else torch.distributed
    if comm is not None:
        allc = comm.gather(cnt.data_ptr(), C, total_channels).astype(np.int64)
        gathered = "qpsk_ber_gather (ncclAllGather behind the C ABI)"
    else:
        allc = shard.gather_counters(cnt, dist).cpu().numpy().astype(np.int64)
        gathered = "all_gather (torch.distributed)" if dist is not None else "single rank"
    samples = total_channels * (ff // 2) * steps
    gbs = 8.0 * samples / world / (ms * 1e-3) / 1e9   # per GPU: 8 B read per complex sample (fused ideal)

    # ---- parity of THIS run: its last step's bits against the oracle fed the same burst sequence -------------------
    n_par = (C if cpu else 0) if parity_channels is None else min(parity_channels, C)
    parity = cpu_leg = None
    if n_par > 0:
        order = [i % K for i in range(warmup + steps)]
        chans = [int(c) for c in np.unique(np.linspace(0, C - 1, n_par).astype(np.int64))]
        idx = torch.tensor(chans, device="cuda")
        sets_host = {k: rx[k][idx].cpu().numpy() for k in sorted(set(order))}
        remap = list(range(len(chans)))
        want, dt, n_samp, cores = _oracle_replay(sets_host, order, remap, fs, rs, alpha, use_fll)
        ref_h = ref[idx].cpu().numpy()

        def gpu_strings(b, n):
            bh, nh = b[idx].cpu().numpy(), n[idx].cpu().numpy()
            return [(bh[j, : nh[j]] + 48).astype(np.uint8).tobytes().decode("ascii") for j in range(len(chans))]

        got = gpu_strings(bits, nb)
        cnt_h = cnt[idx].cpu().numpy().astype(np.int64)
        o_err = [_count_errors(w, ref_h[j]) for j, w in enumerate(want)]
        mode_name = "fast" if fir_mode == Q.FIR_FAST else "exact"
        parity = {"channels_checked": len(chans), "bursts_per_channel": len(order),
                  "what": "bits left by the LAST timed step vs the oracle fed the same burst sequence per channel (state carried)",
                  f"mismatch_{mode_name}": sum(g != w for g, w in zip(got, want)),
                  "oracle_bit_errors": int(sum(o_err)), "gpu_bit_errors": int(cnt_h[:, 0].sum()),
                  "ber_counters_equal": bool(all(int(cnt_h[j, 0]) == o_err[j] for j in range(len(chans)))),
                  "oracle_error_free_channels": int(sum(e == 0 for e in o_err))}
        # the other matched-filter mode on the same sequence (untimed here; "fast_mf" below times it)
        other = Q.FIR_FAST if mode_name == "exact" else Q.FIR_EXACT
        _, _, bits2, nb2, _ = timed(other)
        parity[f"mismatch_{'fast' if other == Q.FIR_FAST else 'exact'}"] = sum(g != w for g, w in zip(gpu_strings(bits2, nb2), want))
        cpu_leg = {"value": n_samp / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port", "seconds": dt,
                   "sample": f"{len(order)} bursts x {len(chans)} channels, C++ restatement of QPSKDeModulator.DeModulate"
                             f"{' with the FLL' if use_fll else ''}, one demodulator per channel, {cores} threads — the run whose bits "
                             "the parity block compares"}
        if dist is not None:
            keys = [k for k in parity if isinstance(parity[k], int) and not isinstance(parity[k], bool) and k != "bursts_per_channel"]
            t = torch.tensor([parity[k] for k in keys] + [int(parity["ber_counters_equal"])], dtype=torch.int64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            for k, v in zip(keys, t.tolist()):
                parity[k] = int(v)
            parity["ber_counters_equal"] = bool(t[-1].item() == world)
            parity["summed_over_ranks"] = world
    fast = None
    if fir_mode is None and cpu:
        ms_f, _, _, _, cnt_f = timed(Q.FIR_FAST)
        fast = {"value": samples / (ms_f * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms_f / steps,
                "bit_errors": int(cnt_f[:, 0].sum().item()),
                "note": "same leg with qpsk_demod_set_fir_mode(QPSK_FIR_FAST): FMA-accumulated matched filter (parity.mismatch_fast "
                        "counts the channels whose bits differ from the oracle's)"}
    return {
        "label": label, "cpu_baseline": cpu_leg, "parity": parity, "fast_mf": fast,
        "mf_mode": "fast" if fir_mode == Q.FIR_FAST else "exact (default: the reference's summation order)",
        "impairments": "two unstable LOs (100 MHz, 1 ppm static error + random-walk drift each), AWGN -40 dBFS, static two-path "
                       "multipath (echo 0.12+0.08j, 3 samples late), regenerated on the device per channel from the counter RNG",
        "workload": f"{total_channels} channels ({C}/GPU) x {ff // 2} cf32 samples per burst, "
                    f"{'FLL -> ' if use_fll else ''}MF(21 taps) -> MM -> Costas -> decode -> TSC strip -> BER; {K} burst sets cycled "
                    f"({set_bytes * K / 1e6:.0f} MB > L2)",
        "channels_total": total_channels, "channels_per_gpu": C,
        "value": samples / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms / steps, "steps": steps,
        "gpu_launches": int(launches),
        "ber": {"channels": int(allc.shape[0]), "bits_per_channel": nref,
                "error_free_channels": int((allc[:, 0] == 0).sum()),
                "bit_errors": int(allc[:, 0].sum()), "bits": int(allc[:, 1].sum()),
                "gathered_with": gathered},
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                     "note": "8 B/sample algorithmic; the serial loops (8 lanes per stream in the FLL, one lane per channel in MM / Costas) are "
                             "dependent-issue-latency bound, not HBM bound (DESIGN.md)"},
    }


def run_modulator(Q, torch, dist, world, rank, stream, steps=3, warmup=3, frames_per_gpu=2048, n_payload=65536,
                  hbm_peak=6461.8, parity=False):
    from qpsk_modulator_demodulator_b200 import shard
    fs, rs = 4000, 1000
    F = frames_per_gpu
    first, _ = shard.channel_range(rank, world, F * world)
    mod = Q.QPSKModulator(fs, rs, 0.35, 10, True, TSC)
    pay = torch.empty((F, n_payload), dtype=torch.uint8, device="cuda")
    Q.fill_bytes_dev(7, first, F, n_payload, pay.data_ptr(), stream)
    ff = mod.frame_floats(n_payload, b"START", b"END")
    stride = ff + (ff & 1)
    out = torch.empty((F, stride), dtype=torch.float32, device="cuda")

    def step():
        mod.modulate_frames_dev(pay.data_ptr(), n_payload, F, b"START", b"END", out.data_ptr(), stride, stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    Q.launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    launches = Q.launch_count()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    samples = world * F * (ff // 2) * steps
    per_gpu_bytes = (8.0 * F * (ff // 2) + F * n_payload * 2.0) * steps   # 8 B written/sample + payload read twice
    gbs = per_gpu_bytes / (ms * 1e-3) / 1e9
    par = None
    if parity:
        # what the last timed step wrote, against the oracle's QPSKModulator.ModulateBytes (fp64 FFT convolution) on the same
        # payloads: the first and the last frame of the batch; north_star's tolerance 1e-5 * max|y|
        import oracle as O
        om = O.QPSKModulator(fs, rs, 0.35, 10, True, TSC)
        worst = 0.0
        for fr in (0, F - 1):
            want = om.ModulateBytes(pay[fr].cpu().numpy().tobytes(), b"START", b"END")
            got = out[fr, :ff].cpu().numpy()
            worst = max(worst, float(np.abs(got - want).max() / np.abs(want).max()))
        par = {"frames_checked": 2, "samples_per_frame": ff // 2, "max_abs_err_over_max_abs": worst, "tolerance": 1e-5,
               "ok": bool(worst <= 1e-5)}
    del out
    return {
        "parity": par,
        "workload": f"{world * F} frames ({F}/GPU) x {n_payload} B payload ({world * F * n_payload / 2**30:.3f} GiB), "
                    f"START|payload|END + 64-bit TSC, differential, sps 4, span 10 (41 taps), alpha .35; "
                    f"{F * (ff // 2) * 8 / 1e9:.1f} GB written per GPU per step",
        "value": samples / (ms * 1e-3) / 1e6, "unit": "Msamples/s (output)", "ms_per_step": ms / steps, "steps": steps,
        "gpu_launches": int(launches),
        "roofline": {"kernel": "mod_shape_kernel", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": gbs / hbm_peak, "algorithmic": "8 B written per output sample + 0.25/sps B read twice"},
    }


def run_stream(Q, blocks=400, n_payload=512, depth=4, cpu_blocks=40):
    """SURVEY §8f-3 leg: one radio stream, one burst per block (4196 cf32 samples), host memory in and payload bytes
    out, wall-clock timed through the public calls: (a) the per-call path DeModulateBytes(block) — H2D, chain, D2H and a
    synchronise per block, what the reference loop does (TB/SDR/ModDemodOverSDR.cs:127-136); (b) the streaming
    front-end push()/poll() with `depth` slots; (c) the same with CS16 ingest; (d) the oracle on one host core."""
    import time
    import oracle as O
    O.build()
    fs = 10_000_000
    rs = fs // 2
    alpha = float(np.float32(0.4))
    S, E = b"S", b"E"
    mod = O.QPSKModulator(fs, rs, alpha, 10, True, TSC)
    tx, rx = O.NCO(100e6, fs, 1, seed=9, stream=0), O.NCO(100e6, fs, 1, seed=9, stream=1)
    uniq = 16
    pays = [O.fill_bytes(2026, 3, 1000 * k, n_payload) for k in range(uniq)]
    bursts = [O.channel_apply(tx, rx, 0, mod.ModulateBytes(p, S, E)) for p in pays]
    n = bursts[0].size
    peak = max(float(np.abs(b).max()) for b in bursts)
    b16 = [np.clip(np.round(b / peak * 30000.0), -32768, 32767).astype(np.int16) for b in bursts]
    out = {"workload": f"{blocks} blocks x {n // 2} cf32 samples (one framed {n_payload}-byte burst each), single stream, "
                       f"host buffers in, payload bytes out, depth {depth}"}

    def rate(dt, nb):
        return nb * (n // 2) / dt / 1e6

    # (a) per call
    d = Q.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC)
    for i in range(8):
        d.DeModulateBytes(bursts[i % uniq], S, E)
    t0 = time.perf_counter()
    ok = 0
    for i in range(blocks):
        ok += bool(d.DeModulateBytes(bursts[i % uniq], S, E))
    out["per_call"] = {"msamples_s": rate(time.perf_counter() - t0, blocks), "frames": ok}
    # (b) streaming, cf32
    d2 = Q.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC)
    st = Q.StreamingDemodulator(d2, S, E, max_block_floats=n, max_payload_bytes=2048, depth=depth)
    for i in range(8):
        st.push(bursts[i % uniq])
    st.drain()
    t0 = time.perf_counter()
    ok = 0
    for i in range(blocks):
        st.push(bursts[i % uniq])
        while True:
            p = st.poll()
            if p is None:
                break
            ok += bool(p)
    ok += sum(bool(p) for p in st.drain())
    out["stream"] = {"msamples_s": rate(time.perf_counter() - t0, blocks), "frames": ok,
                     "h2d_bytes_per_block": n * 4, "d2h_bytes_per_block": 2048 + 8}
    st.close()
    # (c) streaming, CS16
    d3 = Q.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC)
    st = Q.StreamingDemodulator(d3, S, E, max_block_floats=n, max_payload_bytes=2048, depth=depth)
    sc = peak / 30000.0
    for i in range(8):
        st.push_cs16(b16[i % uniq], sc)
    st.drain()
    t0 = time.perf_counter()
    ok = 0
    for i in range(blocks):
        st.push_cs16(b16[i % uniq], sc)
        while True:
            p = st.poll()
            if p is None:
                break
            ok += bool(p)
    ok += sum(bool(p) for p in st.drain())
    out["stream_cs16"] = {"msamples_s": rate(time.perf_counter() - t0, blocks), "frames": ok, "h2d_bytes_per_block": n * 2}
    st.close()
    # (d) oracle, one core
    od = O.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC)
    od.DeModulateBytes(bursts[0], S, E, cap=1 << 16)
    t0 = time.perf_counter()
    ok = 0
    for i in range(cpu_blocks):
        ok += bool(od.DeModulateBytes(bursts[(i + 1) % uniq], S, E, cap=1 << 16))
    out["cpu_oracle_1core"] = {"msamples_s": rate(time.perf_counter() - t0, cpu_blocks), "frames": ok}
    return out


def _median(v):
    v = sorted(v)
    return v[len(v) // 2]


def run_chain_e2e(Q, torch, dist, world, rank, stream, steps=5, channels_per_gpu=2048, use_fll=False, n_payload=512):
    """End-to-end chain leg: what the modem ships.  HOST samples in ([channels][n_floats], 8 B — or 4 B as CS16 — per complex
    sample), payload BYTES out, through the host entry point qpsk_demod_bytes / qpsk_demod_bytes_cs16 (DeModulateBytes): the
    copy of time chunk t+1 runs under the chain on chunk t, the payloads come back in one copy.  Wall-clock per call
    (H2D + FLL? + MF + MM + Costas + decode + TSC strip + framer + D2H inside), median step, max over ranks.  The same call is
    timed on page-locked (qpsk_host_alloc), registered-in-place (qpsk_host_register: a C# float[] pinned by a GCHandle)
    and pageable memory."""
    import time
    from qpsk_modulator_demodulator_b200 import shard
    SM, EM = b"MESSAGE_START", b"MESSAGE_STOP"                   # testAtDataLevel.cs:33-34
    fs = 10_000_000
    rs = fs // 2
    alpha = float(np.float32(0.4))
    C = channels_per_gpu
    first, _ = shard.channel_range(rank, world, C * world)
    seed = 2026
    mod = Q.QPSKModulator(fs, rs, alpha, 10, True, TSC)
    pay = torch.empty((C, n_payload), dtype=torch.uint8, device="cuda")
    Q.fill_bytes_dev(seed, first, C, n_payload, pay.data_ptr(), stream)
    ff = mod.frame_floats(n_payload, SM, EM)
    tx = torch.empty((C, ff), dtype=torch.float32, device="cuda")
    mod.modulate_frames_dev(pay.data_ptr(), n_payload, C, SM, EM, tx.data_ptr(), ff, stream)
    chan = Q.SimChannel(100e6, 100e6, fs, 1, 1, noise_dbfs=-40.0, mode=1, path_gains_iq=(1.0, 0.0, 0.12, 0.08), path_delays=(0, 3),
                        seed=seed, channels=C, first_channel=first)
    rx = torch.empty((C, ff), dtype=torch.float32, device="cuda")
    chan.apply_dev(tx.data_ptr(), ff, ff, rx.data_ptr(), ff, stream)
    torch.cuda.synchronize()
    pay_h = pay.cpu().numpy()
    n = C * ff
    pinned = Q.PinnedBuffer(n)
    torch.from_numpy(pinned.array).copy_(rx.reshape(-1))
    torch.cuda.synchronize()
    pageable = np.array(pinned.array, copy=True)
    registered = np.array(pinned.array, copy=True)
    peak = float(np.abs(pageable).max())
    scale = float(np.float32(peak / 30000.0))
    cs16_pin = Q.PinnedBuffer((n + 1) // 2)                     # n int16 values in n/2 floats of pinned memory
    cs16 = cs16_pin.array.view(np.int16)[:n]
    cs16[:] = np.clip(np.round(pageable / scale), -32768, 32767).astype(np.int16)
    cap = 2 * (n_payload + 64)                                  # = the framer ring bound below: every completed frame fits
    out = np.zeros((C, cap), np.uint8)
    nb = np.zeros(C, np.int64)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def leg(ptr, cs16_scale=None, n_items=ff):
        dem = Q.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC, use_fll=use_fll, channels=C, max_frame_bytes=cap)
        for _ in range(2):
            st = dem.demod_bytes_host_ptr(ptr, n_items, SM, EM, out, nb, cs16_scale)
            assert st == 0, st
        barrier()
        Q.launch_count_reset()
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            st = dem.demod_bytes_host_ptr(ptr, n_items, SM, EM, out, nb, cs16_scale)
            ts.append(time.perf_counter() - t0)
            assert st == 0, st
        launches = Q.launch_count()
        t = torch.tensor([_median(ts)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        good = int(sum(out[c, : nb[c]].tobytes() == pay_h[c].tobytes() for c in range(C)))
        return {"value": world * C * (ff // 2) / float(t.item()) / 1e6, "unit": "Msamples/s", "ms_per_step": 1e3 * float(t.item()),
                "step_ms": [round(1e3 * v, 3) for v in ts], "frames_recovered": good, "gpu_launches_per_step": launches // steps}

    res = {"workload": f"{C * world} channels ({C}/GPU) x {ff // 2} samples per burst, host samples in, payload bytes out "
                       f"({'FLL -> ' if use_fll else ''}MF -> MM -> Costas -> decode -> TSC strip -> framer), impairments as in `chain`",
           "api": "qpsk_demod_bytes / qpsk_demod_bytes_cs16 (DeModulateBytes, host pointers; time-chunk copy pipeline)",
           "timing": "wall clock around the call, median step, max over ranks",
           "h2d_bytes_per_step": C * ff * 4, "d2h_bytes_per_step": C * cap + 8 * C}
    res["pinned"] = leg(pinned.array.ctypes.data)
    with Q.RegisteredArray(registered):
        res["registered"] = leg(registered.ctypes.data)
    res["pageable"] = leg(pageable.ctypes.data)
    res["cs16_pinned"] = dict(leg(cs16.ctypes.data, scale, ff), h2d_bytes_per_step=C * ff * 2)
    res["value"], res["unit"] = res["pinned"]["value"], "Msamples/s"
    pinned.free()
    cs16_pin.free()
    return res


def run_modulator_e2e(Q, torch, dist, world, rank, steps=3, frames_per_gpu=256, n_payload=65536):
    """End-to-end modulator leg: HOST payload bytes in, HOST samples out (qpsk_mod_modulate_frames: ModulateBytes over a batch,
    frame groups pipelined kernel / copy-out).  32*sps output bytes per payload byte: bound by the device-to-host copy."""
    import time
    fs, rs = 4000, 1000
    F = frames_per_gpu
    mod = Q.QPSKModulator(fs, rs, 0.35, 10, True, TSC)
    rng = np.random.default_rng(100 + rank)
    pay = rng.integers(0, 256, (F, n_payload), dtype=np.uint8)
    ff = mod.frame_floats(n_payload, b"START", b"END")
    outp = Q.PinnedBuffer(F * ff)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    mod.ModulateFrames(pay, b"START", b"END", out_ptr=outp.array.ctypes.data, out_stride_floats=ff)
    barrier()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        mod.ModulateFrames(pay, b"START", b"END", out_ptr=outp.array.ctypes.data, out_stride_floats=ff)
        ts.append(time.perf_counter() - t0)
    t = torch.tensor([_median(ts)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    res = {"workload": f"{F * world} frames ({F}/GPU) x {n_payload} B payload, host bytes in, host samples out (sps 4, span 10, TSC, "
                       f"differential); {F * ff * 4 / 1e9:.2f} GB of samples per GPU per step",
           "api": "qpsk_mod_modulate_frames (host pointers, pinned output; 3-slot kernel / copy-out pipeline)",
           "value": world * F * (ff // 2) / dt / 1e6, "unit": "Msamples/s (output)", "ms_per_step": 1e3 * dt,
           "step_ms": [round(1e3 * v, 2) for v in ts], "h2d_bytes_per_step": F * n_payload, "d2h_bytes_per_step": F * ff * 4,
           "d2h_gbs": F * ff * 4 / dt / 1e9, "bound": "pcie (device-to-host)"}
    if world == 1:
        # the same call into a pageable output block (a C# float[] that was not registered): one timed step
        outq = np.empty(F * ff, np.float32)
        outq.fill(0.0)                                          # every page resident before the timing (np.zeros maps lazily)
        mod.ModulateFrames(pay[:8], b"START", b"END", out_ptr=outq.ctypes.data, out_stride_floats=ff)
        t0 = time.perf_counter()
        mod.ModulateFrames(pay, b"START", b"END", out_ptr=outq.ctypes.data, out_stride_floats=ff)
        tq = time.perf_counter() - t0
        res["pageable"] = {"value": F * (ff // 2) / tq / 1e6, "unit": "Msamples/s (output)", "steps": 1,
                           "equal_to_pinned_run": bool(np.array_equal(outq, outp.array[: F * ff]))}
        del outq
    outp.free()
    return res
