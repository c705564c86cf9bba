#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call31.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "pytest-h3" 900 python -m pytest tests/test_h3_gpu.py -m gpu -q
grep -E "^=== |passed|failed|FAILED|Error|assert" $L | tail -20
