#!/bin/bash
# Round-2 GPU call 24: predict_action with the content check off the critical path (speculative replay): tests, timing.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_policy_gpu.py -m gpu -q > gpurun_out/r2c24_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c24_pytest.log
timeout 120 python tools/quick_bench_predict.py > gpurun_out/r2c24_predict.txt 2>&1; grep "predict_action\|graph replay" gpurun_out/r2c24_predict.txt
