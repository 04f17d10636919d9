"""The benchmarked chain configuration, pinned in the GPU suite (round-1 review: the bench ran the default matched-filter
mode with the two-path multipath on, while every parity test of size used QPSK_FIR_EXACT without multipath).

Exactly bench_chain.run_chain's workload — 512-byte random payloads, 64-bit TSC, testAtDataLevel rates, two unstable LOs,
AWGN -40 dBFS, echo 0.12+0.08j three samples late, impairments generated on the device — at 256 channels, with and without
the FLL, in BOTH matched-filter modes: the bits the timed steps leave behind must equal the oracle's, fed the same burst
sequence per channel (state carried over 6 bursts), and the GPU's BER counters must equal a recount from the oracle's bits."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_fll", [False, True])
@pytest.mark.parametrize("mode", ["default", "fast"])
def test_benchmarked_chain_configuration_matches_the_oracle(gpu, orc, use_fll, mode):
    import torch
    import bench_chain
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    r = bench_chain.run_chain(gpu, torch, None, 1, 0, ts.cuda_stream, steps=3, warmup=3, channels_per_gpu=256, use_fll=use_fll,
                              parity_channels=256, fir_mode=(gpu.FIR_FAST if mode == "fast" else None))
    p = r["parity"]
    assert p["channels_checked"] == 256 and p["bursts_per_channel"] == 6
    assert p["mismatch_exact"] == 0, p                          # the default mode: bit-exact is the bar
    assert p["mismatch_fast"] == 0, p                           # FMA-accumulated filter: no decision flips on these bursts either
    assert p["ber_counters_equal"] and p["gpu_bit_errors"] == p["oracle_bit_errors"]
    assert r["mf_mode"].startswith("fast" if mode == "fast" else "exact")
    assert r["ber"]["channels"] == 256 and r["ber"]["error_free_channels"] > 128
    assert "multipath" in r["impairments"]


def test_bench_chain_e2e_legs_recover_the_frames(gpu, orc):
    """The end-to-end legs (host samples in, payload bytes out; cf32 pinned / registered / pageable and CS16) at a small size:
    every variant returns the same number of intact frames, most of the channels'."""
    import torch
    import bench_chain
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    r = bench_chain.run_chain_e2e(gpu, torch, None, 1, 0, ts.cuda_stream, steps=2, channels_per_gpu=256, use_fll=False)
    got = {k: r[k]["frames_recovered"] for k in ("pinned", "registered", "pageable", "cs16_pinned")}
    assert got["pinned"] == got["registered"] == got["pageable"] and got["pinned"] > 200, got
    assert got["cs16_pinned"] > 200, got
    m = bench_chain.run_modulator_e2e(gpu, torch, None, 1, 0, steps=1, frames_per_gpu=16, n_payload=4096)
    assert m["value"] > 0 and m["d2h_bytes_per_step"] > 16 * 4096 * 32
