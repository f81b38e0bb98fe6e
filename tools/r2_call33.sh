#!/bin/bash
# Round-2 GPU call 33: dual Conv3d launch with 12 epilogue warps (setmaxnreg): kernel tests, video parity, A/B vs 8 warps.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "dual" > gpurun_out/r2c33_dual.log 2>&1; echo "dual(ew12) rc=$?"; tail -3 gpurun_out/r2c33_dual.log
timeout 600 python -m pytest tests/test_video_gpu.py -m gpu -q -x > gpurun_out/r2c33_video.log 2>&1; echo "video rc=$?"; grep -v "sampling loop" gpurun_out/r2c33_video.log | tail -3
timeout 900 python tools/ab_forward.py V2A_DUAL_EW=12 V2A_DUAL_EW=8 > gpurun_out/r2c33_ab.txt 2>&1; cat gpurun_out/r2c33_ab.txt
