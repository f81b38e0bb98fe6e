#!/bin/bash
# Round-2 GPU call 21: kernel durations of one predict_action graph replay (ncu, time only) after the round's changes.
mkdir -p gpurun_out
timeout 600 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2c21_predict_launches.csv python tools/profile_predict_target.py > gpurun_out/r2c21_predict.log 2>&1
python tools/launch_shares.py gpurun_out/r2c21_predict_launches.csv 14 > gpurun_out/r2c21_predict_shares.md 2>&1; cat gpurun_out/r2c21_predict_shares.md
