#!/bin/bash
# Round-2 GPU call 39: ncu --set full of the dual Conv3d launches (the kernel the window of call 34 did not contain).
mkdir -p gpurun_out /tmp/ncu
STEPS=1 timeout 900 ncu --clock-control none --set full -k regex:igemm_dual --launch-skip 11 -c 11 -o /tmp/ncu/r2c39_dual_full -f python tools/profile_target.py > gpurun_out/r2c39.log 2>&1
tail -2 gpurun_out/r2c39.log
python tools/ncu_summary.py /tmp/ncu/r2c39_dual_full.ncu-rep > gpurun_out/r2c39_dual_ncu_full_summary.md 2>&1; cut -c1-330 gpurun_out/r2c39_dual_ncu_full_summary.md | head -20
