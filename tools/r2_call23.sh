#!/bin/bash
# Round-2 GPU call 23: small-M kernel with branch-free row chunks + halving reduce: tests, predict timing, durations.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_policy_gpu.py -m gpu -q > gpurun_out/r2c23_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c23_pytest.log
timeout 120 python tools/quick_bench_predict.py > gpurun_out/r2c23_predict.txt 2>&1; grep "predict_action\|graph replay" gpurun_out/r2c23_predict.txt
timeout 600 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2c23_predict_launches.csv python tools/profile_predict_target.py > gpurun_out/r2c23_predict.log 2>&1
python tools/launch_shares.py gpurun_out/r2c23_predict_launches.csv 8 > gpurun_out/r2c23_predict_shares.md 2>&1; cat gpurun_out/r2c23_predict_shares.md
