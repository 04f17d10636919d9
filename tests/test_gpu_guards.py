"""Guard-band and repeatability checks of the device entry points — this pool refuses compute-sanitizer
(profiles/r02_compute_sanitizer_refused.log), so out-of-bounds writes and races are hunted with checks of our own:

  * every device output buffer sits between two guard bands filled with a sentinel; after the call the bands must be
    untouched and every byte the call does not own (row padding between channels) must keep its sentinel;
  * every call is repeated on fresh handles: the hand-rolled pipelines (mbarrier rings, two-warp shared-memory hand-offs,
    cp.async double buffers, two-stream time-chunk pipeline) must give bit-identical results every time — a race shows up
    as run-to-run differences;
  * sizes sit on the awkward edges: one sample short of / past a tile, odd lengths, channel counts that leave partial
    warps and CTAs.
Parity with the oracle is covered by the other test files; here the reference is the call itself."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TSC = "11001010011101100100100110101100" + "01110100111001011010001101101001"
GUARD = 4096          # bytes on each side
SENT = 0xA5


class Guarded:
    """A device buffer of `nbytes` bytes (16-byte aligned) between two sentinel-filled guard bands."""

    def __init__(self, torch, nbytes):
        self.torch = torch
        self.n = int(nbytes)
        self.raw = torch.full((self.n + 2 * GUARD,), SENT, dtype=torch.uint8, device="cuda")
        self.ptr = self.raw.data_ptr() + GUARD

    def body(self):
        return self.raw[GUARD:GUARD + self.n]

    def check(self, what):
        t = self.torch
        assert bool((self.raw[:GUARD] == SENT).all()), f"{what}: wrote before the buffer"
        assert bool((self.raw[GUARD + self.n:] == SENT).all()), f"{what}: wrote past the buffer"

    def f32(self):
        return self.body().view(self.torch.float32)


def _x(torch, gpu, n_floats, seed=1):
    x = torch.empty(n_floats, dtype=torch.float32, device="cuda")
    gpu.fill_uniform_dev(seed, 0, 0, n_floats, x.data_ptr(), 0)
    torch.cuda.synchronize()
    return x


@pytest.mark.parametrize("L", [1, 2559, 2560, 2561, 5121, 40003])
def test_fir_dev_outputs_stay_in_bounds_and_repeat(gpu, L):
    import torch
    for span, sps in ((16, 2), (16, 16)):
        taps = gpu.real_taps_to_iq(gpu.RRCFilter.generateCoefficents(span, 0.35, sps * 1000, 1000))
        for C, pad in ((1, 0), (3, 6)):
            ld = 2 * L + (2 * L) % 4 + pad                       # even float stride, sometimes wider than the row
            x = _x(torch, gpu, C * ld)
            ref = None
            for mode in (gpu.FIR_FAST, gpu.FIR_EXACT):
                outs = []
                for rep in range(3):
                    g = Guarded(torch, C * ld * 4)
                    f = gpu.ComplexFIRFilter(taps, channels=C)
                    f.set_mode(mode)
                    f.filter_dev(x.data_ptr(), g.ptr, 2 * L, ld, ld)
                    torch.cuda.synchronize()
                    g.check(f"fir L={L} C={C} mode={mode}")
                    y = g.f32().view(C, ld)
                    assert bool((y[:, 2 * L:].contiguous().view(torch.uint8) == SENT).all()), "row padding written"
                    outs.append(y[:, : 2 * L].clone())
                assert all(torch.equal(outs[0], o) for o in outs[1:]), (L, C, mode)
            del ref


@pytest.mark.parametrize("dec", [2, 4, 8, 16, 3])
def test_decimate_dev_outputs_stay_in_bounds_and_repeat(gpu, dec):
    import torch
    taps = gpu.real_taps_to_iq(gpu.RRCFilter.generateCoefficents(16, 0.35, 4000, 1000))
    for L in (1, dec - 1, dec, 7 * 256 * dec - 1, 7 * 256 * dec + 1, 7 * 32 * dec + 1, 30011):
        if L < 1:
            continue
        C = 2
        n_out = (L + dec - 1) // dec
        ldo = 2 * n_out + 4
        x = _x(torch, gpu, C * 2 * (L + 1), seed=3)
        outs = []
        for rep in range(3):
            g = Guarded(torch, C * ldo * 4)
            f = gpu.ComplexFIRFilter(taps, channels=C)
            assert f.decimate_dev(x.data_ptr(), 2 * L, dec, g.ptr, 2 * n_out, in_stride=2 * (L + 1), out_stride=ldo) == 2 * n_out
            torch.cuda.synchronize()
            g.check(f"decimate L={L} D={dec}")
            y = g.f32().view(C, ldo)
            assert bool((y[:, 2 * n_out:].contiguous().view(torch.uint8) == SENT).all())
            outs.append(y[:, : 2 * n_out].clone())
        assert all(torch.equal(outs[0], o) for o in outs[1:]), (L, dec)


@pytest.mark.parametrize("use_fll", [False, True])
def test_demod_dev_outputs_stay_in_bounds_and_repeat(gpu, use_fll):
    import torch
    fs, rs = 10_000_000, 5_000_000
    alpha = float(np.float32(0.4))
    for C, n_payload in ((1, 40), (33, 200), (70, 600)):
        mod = gpu.QPSKModulator(fs, rs, alpha, 10, True, TSC)
        pay = torch.empty((C, n_payload), dtype=torch.uint8, device="cuda")
        gpu.fill_bytes_dev(4, 0, C, n_payload, pay.data_ptr(), 0)
        ff = mod.frame_floats(n_payload, b"S", b"E")
        gtx = Guarded(torch, C * ff * 4)
        mod.modulate_frames_dev(pay.data_ptr(), n_payload, C, b"S", b"E", gtx.ptr, ff)
        torch.cuda.synchronize()
        gtx.check("modulator")
        grx = Guarded(torch, C * ff * 4)
        ch = gpu.SimChannel(100e6, 100e6, fs, 1, 1, noise_dbfs=-40.0, mode=1, path_gains_iq=(1.0, 0.0, 0.12, 0.08), path_delays=(0, 3),
                            seed=9, channels=C, first_channel=0)
        ch.apply_dev(gtx.ptr, ff, ff, grx.ptr, ff)
        torch.cuda.synchronize()
        grx.check("channel")
        runs = []
        for rep in range(3):
            dem = gpu.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC, use_fll=use_fll, channels=C, max_frame_bytes=2048)
            cap = dem.bits_bound(ff)
            gb = Guarded(torch, C * cap)
            gn = Guarded(torch, C * 8)
            gp = Guarded(torch, C * 1024)
            gnp = Guarded(torch, C * 8)
            dem.demod_bits_dev(grx.ptr, ff, ff, gb.ptr, cap, gn.ptr)
            dem2 = gpu.QPSKDeModulator(fs, rs, alpha, 10, tsc=TSC, use_fll=use_fll, channels=C, max_frame_bytes=2048)
            dem2.demod_bytes_dev(grx.ptr, ff, ff, b"S", b"E", gp.ptr, 1024, gnp.ptr)
            torch.cuda.synchronize()
            for g, w in ((gb, "bits"), (gn, "bit counts"), (gp, "payload"), (gnp, "payload lengths")):
                g.check(f"demod {w} C={C} fll={use_fll}")
            nb = gn.body().view(torch.int64).clone()
            bits = gb.body().view(C, cap)
            for c in range(C):                                   # nothing written past a channel's own bit count
                assert bool((bits[c, int(nb[c]):] == SENT).all()) or int(nb[c]) == 0, (C, c)
            runs.append((nb, torch.stack([bits[c, : int(nb.min())] for c in range(C)]).clone(), gnp.body().clone()))
        for r in runs[1:]:
            assert torch.equal(runs[0][0], r[0]) and torch.equal(runs[0][1], r[1]) and torch.equal(runs[0][2], r[2])


def test_loop_dev_outputs_stay_in_bounds_and_repeat(gpu):
    import torch
    for C, L in ((1, 257), (5, 1000), (37, 333)):
        ld = 2 * L + 2
        x = _x(torch, gpu, C * ld, seed=6)
        for size in (40, 10, 13):
            outs = []
            for rep in range(3):
                g = Guarded(torch, C * ld * 4)
                f = gpu.FLLBandEdgeFilter(2.0, 0.4, size, 0.01, channels=C)
                f.process_dev(x.data_ptr(), g.ptr, 2 * L, ld, ld)
                torch.cuda.synchronize()
                g.check(f"fll size={size} C={C}")
                y = g.f32().view(C, ld)
                assert bool((y[:, 2 * L:].contiguous().view(torch.uint8) == SENT).all())
                outs.append(y[:, : 2 * L].clone())
            assert all(torch.equal(outs[0], o) for o in outs[1:]), (C, L, size)
