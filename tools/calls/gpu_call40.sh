#!/bin/bash
# refresh of the round-2 profiling evidence on the final kernels
mkdir -p gpurun_out
L=gpurun_out/call40.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "ncu-launches" 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_h3.csv python tools/one_forward.py 512 h3
run "ncu-gemm" 300 ncu --set full --clock-control none --import-source on -k regex:gemm_h3 -s 4 -c 1 -f -o gpurun_out/r02_ncu_gemm_h3_qkv python tools/ncu_gemm_h3.py gemm 2050 3072 1024
run "ncu-fc1" 300 ncu --set full --clock-control none --import-source on -k regex:gemm_h3 -s 4 -c 1 -f -o gpurun_out/r02_ncu_gemm_h3_fc1_gelu python tools/ncu_gemm_h3.py fc1 2050 4096 1024
run "ncu-fc2" 300 ncu --set full --clock-control none --import-source on -k regex:gemm_h3 -s 4 -c 1 -f -o gpurun_out/r02_ncu_gemm_h3_fc2 python tools/ncu_gemm_h3.py gemm 2050 1024 4096
grep -E "^=== " $L
