#!/bin/bash
# Round-2 GPU call 16: pure kernel durations of the B = 1 predict_action graph (ncu, time only) -- how much of the
# 0.655 ms per UNet1D forward is execution and how much is launch / dependency latency.
mkdir -p gpurun_out
timeout 600 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2c16_predict_launches.csv python tools/profile_predict_target.py > gpurun_out/r2c16_predict.log 2>&1
tail -2 gpurun_out/r2c16_predict.log
python tools/launch_shares.py gpurun_out/r2c16_predict_launches.csv 20 > gpurun_out/r2c16_predict_shares.md 2>&1; cat gpurun_out/r2c16_predict_shares.md
