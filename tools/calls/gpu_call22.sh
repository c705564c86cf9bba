#!/bin/bash
mkdir -p gpurun_out
for st in 10 20; do
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 296$st bench.py --gpus 2 --steps $st --warmup 3 --no-raster --no-multiview 2>gpurun_out/n2_$st.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('steps',d['steps'],'value_ms',d['ms_per_step'],'e2e',d['e2e'])"
done
