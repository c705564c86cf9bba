#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call30.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "pytest-raster" 900 python -m pytest tests/test_ops_gpu.py tests/test_renderer_frontend_gpu.py tests/test_renderer_labels_gpu.py -m gpu -q -x -k "raster or splat or render or feature"
grep -E "^=== |passed|failed|FAILED|Error" $L | tail -12
