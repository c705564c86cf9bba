#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
echo "rc=$?"; tail -c 3000 gpurun_out/r02_bench_n2.json; tail -5 gpurun_out/r02_bench_n2.err
timeout -k 10 600 python -m pytest tests -m gpu -q -k "parallel or gather or nccl or dist" 2>&1 | tail -3
