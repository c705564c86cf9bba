#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call13.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
B="python bench.py --steps 20 --warmup 5 --no-multiview --no-raster --no-cpu-baseline"
run "bench-p1c70" 300 $B
SIU3R_HEAD_CLUSTER_CAP=0 run "bench-p1c0" 300 $B
SIU3R_SEG_PRIORITY=0 SIU3R_HEAD_CLUSTER_CAP=0 run "bench-p0c0" 300 $B
SIU3R_HEAD_CLUSTER_CAP=64 run "bench-p1c64" 300 $B
SIU3R_SEG_PRIORITY=0 SIU3R_HEAD_CLUSTER_CAP=70 run "bench-p0c70" 300 $B
SIU3R_PDL=0 TL_ROWS=0 run "timeline" 600 python tools/timeline.py 512 h3 gpurun_out/timeline_h3_b.json
grep -E "^=== " $L | tail -30; grep -o '"metric": "image_pairs[^}]*"ms_per_step": [0-9.]*' $L
