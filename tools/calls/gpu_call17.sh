#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call17.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "pytest-gpu" 1500 python -m pytest tests -m gpu -q --durations=5
run "smoke" 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
SIU3R_BENCH_SHAPES=1 run "bench" 600 python bench.py --steps 20 --warmup 5
cp gpurun_out/call17.log gpurun_out/call17.copy.log
grep -E "^=== |passed|failed|FAILED|Error|smoke ok" $L | tail -30; grep -o '"metric": "image_pairs[^}]*"ms_per_step": [0-9.]*' $L
