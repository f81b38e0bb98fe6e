#!/bin/bash
# Round-2 GPU call 4: the dual-program Conv3d kernel.  Kernel-level test first (own timeout: a pipeline bug traps
# after ~2 s per wait, never hangs), then the network-level goldens, then a same-box A/B of the B = 16 forward.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "dual" > gpurun_out/r2c4_dual.log 2>&1; echo "dual rc=$?"
tail -15 gpurun_out/r2c4_dual.log
timeout 600 python -m pytest tests/test_video_gpu.py -m gpu -q -x > gpurun_out/r2c4_video.log 2>&1; echo "video rc=$?"
grep -v "sampling loop" gpurun_out/r2c4_video.log | tail -8
timeout 400 python tools/ab_forward.py V2A_DUAL=0 V2A_DUAL=1 V2A_DUAL=1,V2A_DUAL_LAG=1 V2A_DUAL=1,V2A_DUAL_LAG=4 > gpurun_out/r2c4_ab.txt 2>&1; cat gpurun_out/r2c4_ab.txt
timeout 200 python tools/quick_bench.py 16 --layers > gpurun_out/r2c4_layers.txt 2>&1; head -22 gpurun_out/r2c4_layers.txt
