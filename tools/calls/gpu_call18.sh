#!/bin/bash
# round-2 profiling evidence: ncu launch list (time + DRAM bytes) of one h3 forward + one raster frame, --set full captures of the top kernels
mkdir -p gpurun_out
L=gpurun_out/call18.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "ncu-launches" 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_h3.csv python tools/one_forward.py 512 h3
run "ncu-gemm" 300 ncu --set full --clock-control none --import-source on -k regex:gemm_h3 -s 4 -c 1 -f -o gpurun_out/r02_ncu_gemm_h3_qkv python tools/ncu_gemm_h3.py gemm 2050 3072 1024
run "ncu-fc2" 300 ncu --set full --clock-control none --import-source on -k regex:gemm_h3 -s 4 -c 1 -f -o gpurun_out/r02_ncu_gemm_h3_fc2 python tools/ncu_gemm_h3.py gemm 2050 1024 4096
run "ncu-conv" 300 ncu --set full --clock-control none --import-source on -k regex:gemm_h3 -s 4 -c 1 -f -o gpurun_out/r02_ncu_gemm_h3_conv256 python tools/ncu_gemm_h3.py conv 256 256 256 256 3
run "ncu-flash" 300 ncu --set full --clock-control none --import-source on -k regex:flash_h3 -s 4 -c 1 -f -o gpurun_out/r02_ncu_flash_h3 python tools/ncu_gemm_h3.py flash 2 16 1025
run "ncu-render" 300 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -f -o gpurun_out/r02_ncu_render python tools/ncu_raster.py
run "ncu-raster-list" 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_raster.csv python tools/ncu_raster.py
grep -E "^=== " $L; ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches*.csv
