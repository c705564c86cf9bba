#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call38.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
B="python bench.py --steps 30 --warmup 5 --no-multiview --no-raster --no-cpu-baseline"
run "c70" 300 $B
SIU3R_HEAD_CLUSTER_CAP=0 run "c0" 300 $B
run "c70b" 300 $B
SIU3R_HEAD_CLUSTER_CAP=0 run "c0b" 300 $B
SIU3R_HEAD_CLUSTER_CAP=72 run "c72" 300 $B
grep -o '"metric": "image_pairs[^}]*"ms_per_step": [0-9.]*' $L | grep -o '"ms_per_step": [0-9.]*'
