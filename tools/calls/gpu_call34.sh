#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 500 python tools/depth_probe.py 2>&1 | tail -3
