#!/bin/bash
# GPU call: h3 kernel unit tests (one process per group: a sticky CUDA error must not poison the rest), model parity, micro-benchmarks.
mkdir -p gpurun_out
L=gpurun_out/call1.log
: > $L
run() {  # name, timeout, command...
  local name=$1 to=$2; shift 2
  echo "=== $name" >> $L
  timeout -k 10 $to "$@" >> $L 2>&1
  echo "=== $name rc=$?" >> $L
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $L 2>&1
for k in split_roundtrip gemm_h3_plain gemm_h3_epilogue gemm_h3_inplace gemm_h3_group conv2d_h3 layernorm_h3 eltwise_resize test_flash_attn_h3 flash_attn_h3_inside; do
  run "unit:$k" 240 python -m pytest tests/test_h3_gpu.py -q -x -k "$k" -s
done
run "model:h3-64" 300 python -m pytest tests/test_model_gpu.py -q -x -k "north_star and h3 and 64" -s
run "model:h3-256" 300 python -m pytest tests/test_model_gpu.py -q -x -k "north_star and h3 and 256" -s
run "model:h3-512" 300 python -m pytest tests/test_model_gpu.py -q -x -k "north_star and h3 and 512" -s
run "bench" 600 python tools/h3_bench.py gemm conv flash model
run "full:h3" 600 python -m pytest tests/test_fulltensor_gpu.py -q -k "h3" -s
run "full:envelope" 600 python -m pytest tests/test_fulltensor_gpu.py -q -k "envelope" -s
grep -E "^=== |passed|failed|error" $L | tail -60
