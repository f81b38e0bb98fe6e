#!/bin/bash
# One GPU call that collects the round's evidence (run under gpurun, one GPU):
#  1. per-layer CUDA-event timings (developer probes, not bench values)
#  2. launch list of 2 denoise steps with duration + DRAM bytes per launch (cheap metrics pass)
#  3. ncu --set full captures of each non-igemm kernel family (video + policy)
mkdir -p gpurun_out
NCU="ncu --clock-control none"
python tools/quick_bench.py 16 --layers > gpurun_out/layers_video.txt 2>&1
python tools/quick_bench_policy.py 256 --layers > gpurun_out/layers_policy.txt 2>&1
STEPS=2 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --csv --log-file gpurun_out/launches_video.csv python tools/profile_target.py > gpurun_out/launches_video.log 2>&1
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --csv --log-file gpurun_out/launches_policy.csv -s 600 -c 600 python tools/quick_bench_policy.py 256 > gpurun_out/launches_policy.log 2>&1
STEPS=1 $NCU --set full --import-source on -k regex:prep_kernel -s 6 -c 3 -f -o gpurun_out/full_prep python tools/profile_target.py > gpurun_out/full_prep.log 2>&1
STEPS=1 $NCU --set full --import-source on -k regex:attention_kernel -s 4 -c 3 -f -o gpurun_out/full_attn python tools/profile_target.py > gpurun_out/full_attn.log 2>&1
STEPS=1 $NCU --set full --import-source on -k regex:"gn_finalize|ddim_step|unet_input_pack|unet_output_head|linear_kernel" -s 10 -c 8 -f -o gpurun_out/full_misc python tools/profile_target.py > gpurun_out/full_misc.log 2>&1
$NCU --set full --import-source on -k regex:"gn_act|colsum|im2col_t|scatter_rows|grad_prep|act_bwd|add_strided|sumsq|adamw_ema|gather_split" -s 300 -c 40 -f -o gpurun_out/full_policy python tools/quick_bench_policy.py 256 > gpurun_out/full_policy.log 2>&1
ls -la gpurun_out
