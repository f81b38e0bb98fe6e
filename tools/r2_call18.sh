#!/bin/bash
# Round-2 GPU call 18: per-(sample, group) GroupNorm forward for small batches + small-M v3: tests, predict timing.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_policy_gpu.py -m gpu -q > gpurun_out/r2c18_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c18_pytest.log
timeout 120 python tools/quick_bench_predict.py > gpurun_out/r2c18_predict.txt 2>&1; tail -12 gpurun_out/r2c18_predict.txt
V2A_POLICY_GN_GROUPS=0 timeout 120 python tools/quick_bench_predict.py > gpurun_out/r2c18_predict_gn0.txt 2>&1; grep "predict_action\|graph" gpurun_out/r2c18_predict_gn0.txt
