#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call36.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
B="python bench.py --steps 20 --warmup 5 --no-multiview --no-raster --no-cpu-baseline"
run "bench-rem1" 300 $B
SIU3R_H3_REM=0 run "bench-rem0" 300 $B
run "bench-rem1b" 300 $B
SIU3R_H3_REM=0 run "bench-rem0b" 300 $B
grep -o '"metric": "image_pairs[^}]*"ms_per_step": [0-9.]*' $L
