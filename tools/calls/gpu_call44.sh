#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call44.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "pytest-gpu" 1500 python -m pytest tests -m gpu -q --durations=5
run "smoke" 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
timeout -k 10 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?" >> $L
grep -E "^=== |passed|failed|FAILED|Error|smoke ok|bench rc" $L | tail -30; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
for k in ('roofline','config3','multiview','tf32_mode','raster','cpu_baseline'):
    v=d.get(k); print(k, json.dumps(v)[:400])
P
