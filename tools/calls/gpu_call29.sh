#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call29.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
B="python bench.py --steps 24 --warmup 6 --no-multiview --no-raster --no-cpu-baseline"
run "slots2" 300 $B
SIU3R_BENCH_SLOTS=3 run "slots3" 300 $B
SIU3R_BENCH_SLOTS=4 run "slots4" 300 $B
grep -E "^=== " $L | tail; grep -o '"metric": "image_pairs[^}]*"ms_per_step": [0-9.]*' $L; grep -o '"e2e": {"value": [0-9.]*' $L
