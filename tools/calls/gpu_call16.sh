#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call16.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
B="python bench.py --steps 20 --warmup 5 --no-multiview --no-raster --no-cpu-baseline"
run "pytest-h3" 900 python -m pytest tests/test_h3_gpu.py -m gpu -q -x
run "flash" 300 python tools/h3_bench.py flash epilogue
run "bench" 300 $B
SIU3R_FUSE_LN=0 run "bench-fuse0" 300 $B
grep -E "^=== |passed|failed|FAILED|Error" $L | tail -30; grep '"kind": "epilogue"\|layernorm_h3\|"kind": "flash"' $L | cut -c1-700; grep -o '"metric": "image_pairs[^}]*"ms_per_step": [0-9.]*' $L
