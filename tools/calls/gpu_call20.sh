#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call20.log
: > $L
echo "=== nproc $(nproc)" >> $L
timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/e2e_probe.py >> $L 2>&1
OMP_NUM_THREADS=8 timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/e2e_probe.py >> $L 2>&1
grep -E "^\{|nproc" $L
