#!/bin/bash
# GPU call 2: whole GPU suite, smoke, ncu launch list (time + DRAM bytes) of one h3 forward, bench.py
mkdir -p gpurun_out
L=gpurun_out/call2.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "pytest-gpu" 900 python -m pytest tests -m gpu -q -x --durations=15
run "smoke" 200 python __graft_entry__.py smoke
run "ncu-launches" 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_h3.csv python tools/one_forward.py 512 h3
run "bench" 600 python bench.py --steps 20 --warmup 5
grep -E "^=== |passed|failed|error|Error" $L | tail -30
tail -c 3000 $L
