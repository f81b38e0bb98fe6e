#!/bin/bash
# Round-2 GPU call 30: per-launch table of the B = 1 video forward.
mkdir -p gpurun_out
timeout 200 python tools/quick_bench.py 1 --layers > gpurun_out/r2c30_layers_b1.txt 2>&1; sed -n 1,6p gpurun_out/r2c30_layers_b1.txt; grep -n "by kind" gpurun_out/r2c30_layers_b1.txt
