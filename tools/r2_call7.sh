#!/bin/bash
# Round-2 GPU call 7: ncu --set full with source correlation of (a) the temporal conv launch (K = 384, N = 128), (b) the
# spatial launch before it, (c) the dual launch of the same layer -- where do the epilogue's microseconds go?
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none --profile-from-start off"
V2A_DUAL=0 WARM=0 timeout 300 $NCU -k regex:igemm_kernel -f -o gpurun_out/r2_prof_two python tools/prof_layers.py > gpurun_out/r2c7_two.log 2>&1; tail -5 gpurun_out/r2c7_two.log
V2A_DUAL=1 WARM=0 IGEMMS=2 timeout 300 $NCU -k regex:igemm_dual -c 1 -f -o gpurun_out/r2_prof_dual python tools/prof_layers.py > gpurun_out/r2c7_dual.log 2>&1; tail -5 gpurun_out/r2c7_dual.log
ls -la gpurun_out/*.ncu-rep
