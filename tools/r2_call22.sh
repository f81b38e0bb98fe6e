#!/bin/bash
# Round-2 GPU call 22: ncu --set full of the 16-row small-M launches (18 us each for 1.3 MB of weights: why?)
mkdir -p gpurun_out
timeout 600 ncu --clock-control none --profile-from-start off --set full --import-source on -k regex:igemm_smallm_kernel -s 4 -c 4 -o gpurun_out/r2c22_smallm -f python tools/profile_predict_target.py > gpurun_out/r2c22.log 2>&1
tail -3 gpurun_out/r2c22.log; ls -la gpurun_out/r2c22_smallm.ncu-rep
