#!/bin/bash
# One gpurun call of round 2: GPU tests, the default bench, the CPU arm, then the sanitizer passes.
# usage (from the repo root on the GPU box): bash tools/gpu_round.sh <tag> [steps...]
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(); import qpsk_modulator_demodulator_b200 as Q; print('build id', Q._native.build_id())" > $OUT/${TAG}_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a $OUT/${TAG}_tests.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?" | tee -a $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench.err
