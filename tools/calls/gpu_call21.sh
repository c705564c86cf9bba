#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call21.log
: > $L
for st in 10 20 40; do
timeout -k 10 300 python bench.py --steps $st --warmup 3 --no-multiview --no-raster --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('steps',d['steps'],'value_ms',d['ms_per_step'],'e2e_ms',d['e2e']['ms_per_step'])" >> $L
done
cat $L
