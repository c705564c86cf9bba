#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
echo "rc=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['e2e']['ms_per_step_per_rank'])
print('config3',d.get('config3')); print('allgather',d.get('allgather'))
P
tail -3 gpurun_out/r02_bench_n2.err
