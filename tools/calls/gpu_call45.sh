#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call45.log
: > $L
run() { local name=$1 to=$2; shift 2; echo "=== $name" >> $L; timeout -k 10 $to "$@" >> $L 2>&1; echo "=== $name rc=$?" >> $L; }
run "pytest-raster" 900 python -m pytest tests/test_ops_gpu.py tests/test_renderer_frontend_gpu.py -m gpu -q -x -k "raster or splat or render or feature"
run "sweep" 500 python tools/raster_sweep.py --quick --out gpurun_out/r02_raster_sweep_quick.json
grep -E "^=== |passed|failed|FAILED|Error" $L | tail -8
python - <<'P'
import json
d=json.load(open('gpurun_out/r02_raster_sweep_quick.json'))
for r in d['rows']:
    print(r['H'],r['W'],r['G'],r['splats'][:5], round(r['fps_render_cuda']), round(r['ms_render_cuda'],3), round(r['hbm_frac'],3))
P
