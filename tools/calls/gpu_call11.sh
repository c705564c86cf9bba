#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call11.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
TL_ROWS=0 run "timeline" 600 python tools/timeline.py 512 h3 gpurun_out/timeline_h3.json
tail -5 $L
