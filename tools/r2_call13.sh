#!/bin/bash
# Round-2 GPU call 13 (one GPU): validation of the final default configuration -- full GPU suite, smoke(), ncu launch
# list of one video denoise step (shares + igemm DRAM traffic), default bench line.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c13_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c13_pytest.log
tail -4 gpurun_out/r2c13_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c13_smoke.log 2>&1; tail -2 gpurun_out/r2c13_smoke.log
STEPS=1 timeout 500 $NCU --metrics $M --csv --log-file gpurun_out/r2c13_launches_video.csv python tools/profile_target.py > gpurun_out/r2c13_launches_video.log 2>&1
python tools/launch_shares.py gpurun_out/r2c13_launches_video.csv 16 > gpurun_out/r2c13_shares_video.md 2>&1; tail -3 gpurun_out/r2c13_shares_video.md
python tools/igemm_traffic.py gpurun_out/r2c13_launches_video.csv gpurun_out/r2c13_igemm_traffic.json "gpurun_out/r2c13_launches_video.csv = ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active... over STEPS=1 python tools/profile_target.py (B=16; two executions of the denoise step: eager warm-up + the run); every igemm* launch (dual launches count once); summary in profiles/r2_launch_shares_video_final.md"
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2c13_bench.json 2> gpurun_out/r2c13_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2c13_bench.json
du -sh gpurun_out
