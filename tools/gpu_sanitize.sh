#!/bin/bash
# compute-sanitizer over tools/sanitize_slice.py: memcheck (full slice), racecheck / synccheck / initcheck (small slice).
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
python tools/sanitize_slice.py small > $OUT/${TAG}_sanitize_plain.log 2>&1; echo "plain rc=$?" >> $OUT/${TAG}_sanitize_plain.log
timeout 1200 $CS --tool memcheck --leak-check no --print-limit 50 python tools/sanitize_slice.py > $OUT/${TAG}_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> $OUT/${TAG}_sanitize_memcheck.log
timeout 1500 $CS --tool racecheck --racecheck-report all --print-limit 50 python tools/sanitize_slice.py small > $OUT/${TAG}_sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> $OUT/${TAG}_sanitize_racecheck.log
timeout 900 $CS --tool synccheck --print-limit 50 python tools/sanitize_slice.py small > $OUT/${TAG}_sanitize_synccheck.log 2>&1; echo "synccheck rc=$?" >> $OUT/${TAG}_sanitize_synccheck.log
tail -n 4 $OUT/${TAG}_sanitize_*.log
