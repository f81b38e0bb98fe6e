#!/bin/bash
# Round-2 GPU call 2 (one GPU): full GPU suite (no -x: every failure), predict_action breakdown, policy / encoder
# per-layer tables, ncu launch lists (video denoise step, policy optimisation step) for profiles/.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c2_pytest.log
grep -E "passed|failed|^FAILED|rel-L2|trajectory|forward rel|fast" gpurun_out/r2c2_pytest.log | tail -40
timeout 120 python tools/quick_bench_predict.py > gpurun_out/r2c2_predict.txt 2>&1; cat gpurun_out/r2c2_predict.txt | tail -12
timeout 200 python tools/quick_bench_policy.py 256 --layers > gpurun_out/r2c2_layers_policy.txt 2>&1; head -6 gpurun_out/r2c2_layers_policy.txt
timeout 200 python tools/quick_bench_encoder.py 256 --layers > gpurun_out/r2c2_layers_encoder.txt 2>&1; head -4 gpurun_out/r2c2_layers_encoder.txt
STEPS=1 timeout 400 $NCU --metrics $M --csv --log-file gpurun_out/r2c2_launches_video.csv python tools/profile_target.py > gpurun_out/r2c2_launches_video.log 2>&1
STEPS=1 timeout 400 $NCU --profile-from-start off --metrics $M --csv --log-file gpurun_out/r2c2_launches_policy.csv python tools/profile_policy_target.py > gpurun_out/r2c2_launches_policy.log 2>&1
python tools/launch_shares.py gpurun_out/r2c2_launches_video.csv 16 > gpurun_out/r2c2_shares_video.md 2>&1; tail -3 gpurun_out/r2c2_shares_video.md
python tools/launch_shares.py gpurun_out/r2c2_launches_policy.csv 30 > gpurun_out/r2c2_shares_policy.md 2>&1; tail -3 gpurun_out/r2c2_shares_policy.md
du -sh gpurun_out
