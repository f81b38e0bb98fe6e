#!/bin/bash
# policy-path evidence (run under gpurun, one GPU): launch shares of one steady-state compute_loss step and
# ncu --set full captures of the encoder / weight-gradient kernel families (gpurun_out must stay < 64 MiB)
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
STEPS=1 timeout 400 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --csv --log-file gpurun_out/launches_policy.csv python tools/profile_policy_target.py > gpurun_out/launches_policy.log 2>&1
STEPS=1 timeout 300 $NCU --set full -k regex:wgrad_kernel -c 8 -f -o gpurun_out/full_wgrad python tools/profile_policy_target.py > gpurun_out/full_wgrad.log 2>&1
STEPS=1 timeout 300 $NCU --set full -k regex:"enc_" -c 24 -f -o gpurun_out/full_enc python tools/profile_policy_target.py > gpurun_out/full_enc.log 2>&1
for f in launches_policy full_wgrad full_enc; do tail -n 2 gpurun_out/$f.log; done
du -sh gpurun_out
