#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call4.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "pytest-sel" 900 python -m pytest tests -m gpu -q --durations=5 -k "raster or populated or h3 or fulltensor or renderer or postprocess or rect or north_star or tf32"
run "epilogue-bench" 300 python tools/h3_bench.py epilogue
run "ncu-gemm" 300 ncu --set full --clock-control none --import-source on -k regex:gemm_h3 -s 4 -c 1 -o gpurun_out/r02_ncu_gemm_h3_qkv python tools/ncu_gemm_h3.py gemm 2050 3072 1024
run "ncu-conv" 300 ncu --set full --clock-control none --import-source on -k regex:gemm_h3 -s 4 -c 1 -o gpurun_out/r02_ncu_gemm_h3_conv64 python tools/ncu_gemm_h3.py conv 64 64 256 256 3
run "ncu-flash" 300 ncu --set full --clock-control none --import-source on -k regex:flash_h3 -s 4 -c 1 -o gpurun_out/r02_ncu_flash_h3 python tools/ncu_gemm_h3.py flash 2 16 1025
grep -E "^=== |passed|failed|FAILED" $L | tail -40
