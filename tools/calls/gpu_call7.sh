#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call7.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "pytest-h3" 600 python -m pytest tests/test_h3_gpu.py -q -x
run "limits" 300 python tools/h3_bench.py limits epilogue flash conv
run "model" 300 python tools/h3_bench.py model
grep -E "^=== |passed|failed|FAILED" $L | tail; grep '"kind": "limits"\|"kind": "epilogue"\|"kind": "flash"\|"kind": "model"\|"kind": "conv"' $L | cut -c1-420
