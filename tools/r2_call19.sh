#!/bin/bash
# Round-2 GPU call 19: predict_action with (a) the residual 1x1 conv on the side lane, (b) programmatic dependent launch
# for the small-M and GroupNorm kernels: tests, timing with PDL on / off.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_policy_gpu.py -m gpu -q > gpurun_out/r2c19_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c19_pytest.log
for v in 1 0 1 0; do
  echo "V2A_PDL=$v"; V2A_PDL=$v timeout 120 python tools/quick_bench_predict.py > gpurun_out/r2c19_predict_pdl$v.txt 2>&1; grep "predict_action\|graph replay" gpurun_out/r2c19_predict_pdl$v.txt
done
