set -u
OUT=gpurun_out
python -m pytest tests/test_gpu_fir_split.py tests/test_gpu_fir.py tests/test_gpu_decimate.py tests/test_gpu_boundary.py -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-chain --no-decimate"
for nt in 256; do $B > $OUT/s3_$nt.json 2>$OUT/s3_$nt.err; python - <<PY
import json
d=json.loads(open("$OUT/s3_$nt.json").read().strip().splitlines()[-1])
print($nt, d["value"], d["ms_per_step"], [(r["taps"], round(r["ms"],3), r.get("kernel")) for r in d["roofline_by_taps"]])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_split2 -s 9 -c 3 -f -o $OUT/s3_prof $B --steps 2 > $OUT/s3_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py $OUT/s3_prof.ncu-rep $OUT/s3_ncu_full_split_summary.csv
python tools/ncu_src.py $OUT/s3_prof.ncu-rep 30 > $OUT/s3_ncu_src_split.txt 2>&1
rm -f $OUT/s3_prof.ncu-rep
