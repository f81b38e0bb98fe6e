#!/bin/bash
# Round-2 GPU call 34: ncu --set full of the final tree's dominant kernels in one B = 16 denoise step: a window of
# igemm / dual launches in the steady-state forward (skipping the eager warm-up forward); summarised on the box
# (the report itself exceeds the 64 MiB that travel back).
mkdir -p gpurun_out /tmp/ncu
STEPS=1 timeout 900 ncu --clock-control none --set full -k regex:igemm --launch-skip 170 -c 30 -o /tmp/ncu/r2c34_igemm_full -f python tools/profile_target.py > gpurun_out/r2c34.log 2>&1
tail -2 gpurun_out/r2c34.log
python tools/ncu_summary.py /tmp/ncu/r2c34_igemm_full.ncu-rep > gpurun_out/r2c34_igemm_ncu_full_summary.md 2>&1; head -12 gpurun_out/r2c34_igemm_ncu_full_summary.md | cut -c1-400
