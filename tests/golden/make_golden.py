#!/usr/bin/env python
"""Generates tests/golden/golden_r01.npz from the CPU oracle with fixed seeds.

The reference has no golden vectors and cannot run here (C#, no .NET), so these are the ORACLE's
outputs (parity unpinned, see oracle/qpsk_oracle.cpp): they freeze the restatement so that neither
the oracle nor the CUDA path can drift unnoticed.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

TSC = "11001010011101100100100110101100" + "01110100111001011010001101101001"
TEXT = "The Quick Brown fox jump yes yes man good!"
A04 = float(np.float32(0.4))


def main():
    O.build()
    g = {}
    # a1: taps
    g["rrc_10_04_sps2"] = O.RRCFilter.generateCoefficents(10, A04, 10_000_000, 5_000_000)
    g["rrc_16_035_sps4"] = O.RRCFilter.generateCoefficents(16, 0.35, 4000, 1000)
    g["rrc_4_025_sps8"] = O.RRCFilter.generateCoefficents(4, 0.25, 8000, 1000)
    lo, up = O.FLLBandEdgeFilter(2.0, A04, 40, 1e-4).taps()
    g["fll_lower_sps2_a04_40"], g["fll_upper_sps2_a04_40"] = lo, up
    # a3-a5: FIR
    x = O.fill_uniform(1, 0, 0, 2 * 700)
    g["fir_x"] = x
    t65 = O.real_taps_to_iq(g["rrc_16_035_sps4"])
    g["fir_stream_65"] = O.ComplexFIRFilter(t65).Filter(x)
    g["fir_fft_65"] = O.ComplexFIRFilter(t65).fftFilter(x)
    rng = np.random.default_rng(1)
    tc = (rng.standard_normal(2 * 40) / 6).astype(np.float32)
    g["fir_taps_c40"] = tc
    g["fir_stream_c40"] = O.ComplexFIRFilter(tc).Filter(x)
    # a6: modulator, config 1
    fs = 10_000_000
    mod = O.QPSKModulator(fs, fs // 2, A04, 10, tsc=TSC)
    tx = mod.ModulateTextUtf8(TEXT, "MESSAGE_START", "MESSAGE_STOP")
    g["mod_datalevel"] = tx
    g["mod_noshape_sps4"] = O.QPSKModulator(4000, 1000, 0.35, 6).Modulate("0001111000110110", False)
    # channel + a11/a12: four consecutive bursts through one demodulator
    ntx = O.NCO(100e6, fs, 1, seed=7, stream=0)
    nrx = O.NCO(100e6, fs, 1, seed=7, stream=1)
    dem = O.QPSKDeModulator(fs, fs // 2, A04, 10, tsc=TSC)
    dem2 = O.QPSKDeModulator(fs, fs // 2, A04, 10, tsc=TSC)
    bursts, bits, texts, const = [], [], [], []
    for _ in range(4):
        y = O.channel_apply(ntx, nrx, 0, tx)
        bursts.append(y)
        bits.append(dem.DeModulate(y))
        texts.append(dem2.DeModulateTextUtf8(y, "MESSAGE_START", "MESSAGE_STOP"))
    g["chan_bursts"] = np.stack(bursts)
    g["demod_bits"] = np.array(bits)
    g["demod_texts"] = np.array(texts)
    dem3 = O.QPSKDeModulator(fs, fs // 2, A04, 10, tsc=TSC)
    g["demod_constellation_b0"] = dem3.deModulateConstellation(bursts[0])
    # a8-a10: loops on a short impaired sps-4 burst
    m4 = O.QPSKModulator(4000, 1000, 0.35, 10)
    b4 = "".join(np.random.default_rng(2).choice(["0", "1"], 600))
    s4 = m4.Modulate(b4)
    c_tx = O.NCO(1e6, 4000.0, 20, 0.3, seed=9, stream=0)
    c_rx = O.NCO(1e6, 4000.0, 10, 0.1, seed=9, stream=1)
    nz = O.noise_iq(-35.0, s4.size // 2, 9, 2)
    y4 = O.channel_apply(c_tx, c_rx, 1, s4, nz)
    g["loops_in"] = y4
    g["fll_out"] = O.FLLBandEdgeFilter(4.0, 0.35, 40, 0.01).Process(y4)
    mf = O.ComplexFIRFilter(O.real_taps_to_iq(O.RRCFilter.generateCoefficents(10, 0.35, 4000, 1000))).Filter(y4)
    kp, ki = O.mm_gains_from_bw(0.002)
    sym = O.MuellerMuller(4.0, kp, ki).Process(mf)
    g["mm_out"] = sym
    g["costas_out"] = O.CostasLoopQpsk(1000.0, 1000.0 / 120.0).Process(sym)
    d4 = O.QPSKDeModulator(4000, 1000, 0.35, 10, 0.002, 120.0, float(np.float32(0.01)), use_fll=True)
    g["chain_fll_bits"] = np.array(d4.DeModulate(y4))
    # RNG + generators
    g["rng_u64"] = np.array([O.rng_u64(1, s, c) for s in range(3) for c in range(4)], np.uint64)
    g["payload_bytes"] = np.frombuffer(O.fill_bytes(21, 4 * 5 + 3, 0, 64), np.uint8)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_r01.npz")
    np.savez_compressed(out, **g)
    print(out, os.path.getsize(out), "bytes;", len(g), "arrays")


if __name__ == "__main__":
    main()
