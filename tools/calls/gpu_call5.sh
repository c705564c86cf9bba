#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call5.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "limits" 300 python tools/h3_bench.py limits epilogue
run "pytest-sel" 600 python -m pytest tests -m gpu -q -k "rect or model_rejects or pair_pipeline or batch2"
grep -E "^=== |passed|failed|FAILED" $L | tail; grep '"kind": "limits"' $L
