#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call9.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "pytest-gpu" 1200 python -m pytest tests -m gpu -q --durations=5
run "epilogue" 300 python tools/h3_bench.py epilogue
SIU3R_BENCH_SHAPES=1 run "bench" 600 python bench.py --steps 20 --warmup 5
SIU3R_BENCH_SLOTS=3 run "bench-3slots" 300 python bench.py --steps 20 --warmup 5 --no-multiview --no-raster --no-cpu-baseline
grep -E "^=== |passed|failed|FAILED" $L | tail -30; grep '"kind": "epilogue"' $L | cut -c1-400; grep -o '"metric": "image_pairs[^}]*"ms_per_step": [0-9.]*' $L
