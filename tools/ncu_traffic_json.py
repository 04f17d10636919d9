"""profiles/<tag>_ncu_traffic.json from the condensed ncu summaries (what bench.py reads back as `roofline.traffic` and
`ncu_fma_pipe_active_pct`).  usage: python tools/ncu_traffic_json.py [tag]"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
P = os.path.join(ROOT, "profiles")


def rows(name):
    r = list(csv.reader(open(os.path.join(P, name))))
    return [dict(zip(r[0], x)) for x in r[2:]]


def entry(row, split_at="("):
    rd, wr = float(row["dram__bytes_read.sum"]) * 1e9, float(row["dram__bytes_write.sum"]) * 1e9
    return {"kernel": row["Kernel Name"].split(split_at)[0].strip(), "block": row["Block Size"],
            "registers": int(row["launch__registers_per_thread"]),
            "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr,
            "fma_pipe_active_pct": float(row["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]),
            "issue_active_pct": float(row["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
            "ncu_ms": float(row["gpu__time_duration.sum"])}


out = {"source": f"profiles/{tag}_ncu_full_fir_summary.csv (ncu --set full --clock-control none --import-source on "
                 "-k regex:\"fir_tma|fir_split2\" -s 12 -c 4 on `bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-chain "
                 "--no-decimate`: the four launches of the first timed step, 2^28 samples each)",
       "log2_samples": 28, "per_taps": {}}
for taps, row in zip((33, 65, 129, 257), rows(f"{tag}_ncu_full_fir_summary.csv")):
    out["per_taps"][str(taps)] = entry(row)
out["kernel"] = " / ".join(sorted({v["kernel"] for v in out["per_taps"].values()}))
dec = {}
for (t, D), row in zip(((33, 2), (65, 2), (129, 2), (65, 4), (129, 8), (257, 16)), rows(f"{tag}_ncu_full_decim_summary.csv")):
    e = entry(row)
    e["ctas_per_sm"] = int(float(row["launch__occupancy_limit_registers"]))
    e["algorithmic"] = 2 ** 28 * (8 + 8 / D)
    dec[f"{t}/{D}"] = e
out["decimator"] = {"source": f"profiles/{tag}_ncu_full_decim_summary.csv (ncu --set full on tools/decim_probe.py: one launch per "
                              "taps/D case, 2^28 input samples)", "per_case": dec}
json.dump(out, open(os.path.join(P, f"{tag}_ncu_traffic.json"), "w"), indent=1)
for k, v in out["per_taps"].items():
    print(k, v["kernel"], v["ncu_ms"], round(v["traffic"] / (16 * 2 ** 28), 3), v["fma_pipe_active_pct"])
