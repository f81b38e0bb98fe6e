#!/bin/bash
# Round-2 GPU call 5: dual Conv3d launch after the dependency-check fix (A/B), the small-M backend, predict_action.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "dual or small_m" > gpurun_out/r2c5_ops.log 2>&1; echo "ops rc=$?"
tail -12 gpurun_out/r2c5_ops.log
timeout 400 python tools/ab_forward.py V2A_DUAL=0 V2A_DUAL=1 > gpurun_out/r2c5_ab.txt 2>&1; cat gpurun_out/r2c5_ab.txt
timeout 200 python tools/quick_bench.py 16 --layers > gpurun_out/r2c5_layers.txt 2>&1; sed -n 2,4p gpurun_out/r2c5_layers.txt; grep "K   1536\|K   2688\|K   3840\|K   1920" gpurun_out/r2c5_layers.txt | head
timeout 300 python -m pytest tests/test_policy_gpu.py tests/test_video_gpu.py -m gpu -q -x > gpurun_out/r2c5_nets.log 2>&1; echo "nets rc=$?"
grep -v "sampling loop" gpurun_out/r2c5_nets.log | tail -6
timeout 120 python tools/quick_bench_predict.py > gpurun_out/r2c5_predict.txt 2>&1; cat gpurun_out/r2c5_predict.txt | tail -10
