#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call3.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "pytest-gpu" 1200 python -m pytest tests -m gpu -q --durations=10
SIU3R_BENCH_SHAPES=1 run "bench-shapes" 600 python bench.py --steps 10 --warmup 3 --no-multiview --no-raster --no-cpu-baseline
run "stages" 300 python tools/stage_times.py --precision h3
run "stages-serial" 300 python tools/stage_times.py --precision h3 --serial
grep -E "^=== |passed|failed|FAILED" $L | tail -40
