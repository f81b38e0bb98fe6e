#!/bin/bash
# Round-2 GPU call 12: dual Conv3d launch with four 128-column accumulators (unfused MMAs): kernel test, A/B, layers.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "dual" > gpurun_out/r2c12_dual.log 2>&1; echo "dual(acc4) rc=$?"; tail -4 gpurun_out/r2c12_dual.log
V2A_DUAL_ACC4=0 timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "dual" > gpurun_out/r2c12_dual0.log 2>&1; echo "dual(acc2) rc=$?"; tail -2 gpurun_out/r2c12_dual0.log
timeout 600 python -m pytest tests/test_video_gpu.py -m gpu -q -x > gpurun_out/r2c12_video.log 2>&1; echo "video rc=$?"; grep -v "sampling loop" gpurun_out/r2c12_video.log | tail -3
timeout 600 python tools/ab_forward.py V2A_DUAL_ACC4=0 V2A_DUAL_ACC4=1 V2A_DUAL=0 > gpurun_out/r2c12_ab.txt 2>&1; cat gpurun_out/r2c12_ab.txt
timeout 200 python tools/quick_bench.py 16 --layers > gpurun_out/r2c12_layers.txt 2>&1; sed -n 2,4p gpurun_out/r2c12_layers.txt; grep "cout   128" gpurun_out/r2c12_layers.txt | grep "#"
