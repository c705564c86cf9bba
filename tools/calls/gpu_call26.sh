#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call26.log
: > $L
timeout -k 10 300 python tools/h3_bench.py limits >> $L 2>&1
grep '"kind": "limits"' $L
