#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
echo "rc=$?"; python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'])
    print('config3',d.get('config3')); print('allgather',d.get('allgather')); print('clocks',d.get('clocks'))
except Exception as e:
    print('parse failed',e)
P
tail -5 gpurun_out/r02_bench_n8.err; nproc; free -g | head -2
