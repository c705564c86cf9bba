#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 3 --no-raster > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err
echo "rc=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_n4.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['e2e']['d2h_pinned_GBs'])
print('config3',d.get('config3',{}).get('value')); print('allgather',d.get('allgather',{}).get('busbw_GBs'))
P
