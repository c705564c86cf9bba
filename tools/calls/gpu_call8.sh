#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call8.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "pytest-gpu" 1200 python -m pytest tests -m gpu -q --durations=8
run "smoke" 200 python __graft_entry__.py smoke
run "bench" 600 python bench.py --steps 20 --warmup 5
run "stages" 200 python tools/stage_times.py --precision h3
grep -E "^=== |passed|failed|FAILED" $L | tail -30; grep -o '"metric": "image_pairs[^}]*"ms_per_step": [0-9.]*' $L | head -2; grep "t = " $L
