#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call28.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
B="python bench.py --steps 20 --warmup 5 --no-multiview --no-raster --no-cpu-baseline"
run "pytest-h3" 900 python -m pytest tests/test_h3_gpu.py -m gpu -q -x
run "bench-prew1" 300 $B
SIU3R_H3_PREW=0 run "bench-prew0" 300 $B
run "bench-prew1b" 300 $B
run "pytest-model" 900 python -m pytest tests/test_model_gpu.py tests/test_multiview_gpu.py -m gpu -q -x
grep -E "^=== |passed|failed|FAILED|Error" $L | tail -30; grep -o '"metric": "image_pairs[^}]*"ms_per_step": [0-9.]*' $L
