#!/bin/bash
# Round-2 GPU call 17: small-M backend v3 (four channels per warp, host-side offset table, weight prefetch before the
# barrier): kernel tests, policy parity tests, predict_action timing, kernel durations of one predict_action graph.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_policy_gpu.py -m gpu -q -k "small_m or predict or policy" > gpurun_out/r2c17_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c17_pytest.log
timeout 120 python tools/quick_bench_predict.py > gpurun_out/r2c17_predict.txt 2>&1; tail -12 gpurun_out/r2c17_predict.txt
timeout 600 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2c17_predict_launches.csv python tools/profile_predict_target.py > gpurun_out/r2c17_predict.log 2>&1
python tools/launch_shares.py gpurun_out/r2c17_predict_launches.csv 8 > gpurun_out/r2c17_predict_shares.md 2>&1; cat gpurun_out/r2c17_predict_shares.md
