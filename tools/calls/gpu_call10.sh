#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call10.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "pytest-h3" 900 python -m pytest tests/test_h3_gpu.py tests/test_rect_gpu.py tests/test_model_gpu.py tests/test_fulltensor_gpu.py -m gpu -q -x
run "bench-pdl1" 300 python bench.py --steps 20 --warmup 5 --no-multiview --no-raster --no-cpu-baseline
SIU3R_PDL=0 run "bench-pdl0" 300 python bench.py --steps 20 --warmup 5 --no-multiview --no-raster --no-cpu-baseline
grep -E "^=== |passed|failed|FAILED" $L | tail -30; grep -o '"metric": "image_pairs[^}]*"ms_per_step": [0-9.]*' $L
