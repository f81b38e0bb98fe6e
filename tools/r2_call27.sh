#!/bin/bash
# Round-2 GPU call 27: second same-box A/B of programmatic dependent launch on the tensor-core kernels (final tree).
mkdir -p gpurun_out
timeout 900 python tools/ab_forward.py V2A_PDL=1 V2A_PDL=0 V2A_PDL=1 > gpurun_out/r2c27_ab.txt 2>&1; cat gpurun_out/r2c27_ab.txt
timeout 900 python tools/ab_forward.py V2A_PDL=0 V2A_PDL=1 > gpurun_out/r2c27_ab2.txt 2>&1; cat gpurun_out/r2c27_ab2.txt
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
