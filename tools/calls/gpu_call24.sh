#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call24.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "ncu-fc1" 300 ncu --set full --clock-control none --import-source on -k regex:gemm_h3 -s 4 -c 1 -f -o gpurun_out/r02_ncu_gemm_h3_fc1_gelu python tools/ncu_gemm_h3.py fc1 2050 4096 1024
run "ncu-fc1-plain" 300 ncu --set full --clock-control none --import-source on -k regex:gemm_h3 -s 4 -c 1 -f -o gpurun_out/r02_ncu_gemm_h3_fc1_plain python tools/ncu_gemm_h3.py gemm 2050 4096 1024
grep -E "^=== " $L
