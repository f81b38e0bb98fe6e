#!/bin/bash
# Round-2 GPU call 20: predict_action with the timestep-MLP table: policy tests, timing.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_policy_gpu.py -m gpu -q > gpurun_out/r2c20_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c20_pytest.log
timeout 120 python tools/quick_bench_predict.py > gpurun_out/r2c20_predict.txt 2>&1; grep "predict_action\|graph replay" gpurun_out/r2c20_predict.txt
