#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 500 python tools/raster_sweep.py --out gpurun_out/r02_raster_sweep.json 2>&1 | tail -22 | cut -c1-330
