#!/bin/bash
# Round-2 GPU call 14: the four small-kernel changes (vectorised fingerprint, 32-bit input-pack indexing, tiled stencil9,
# SiLU hoisted out of the wide emb linear): ops + video parity tests, forward time, per-launch table, predict_action.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_video_gpu.py -m gpu -q > gpurun_out/r2c14_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "sampling loop" gpurun_out/r2c14_pytest.log | tail -4
timeout 300 python tools/ab_forward.py V2A_DUAL=1 > gpurun_out/r2c14_ab.txt 2>&1; cat gpurun_out/r2c14_ab.txt
timeout 200 python tools/quick_bench.py 16 --layers > gpurun_out/r2c14_layers.txt 2>&1; sed -n 1,4p gpurun_out/r2c14_layers.txt; grep -n "by kind" gpurun_out/r2c14_layers.txt
timeout 120 python tools/quick_bench_predict.py > gpurun_out/r2c14_predict.txt 2>&1; tail -12 gpurun_out/r2c14_predict.txt
