#!/bin/bash
# Two-GPU checks for the next round (gpurun --gpus 2 --timeout 300 -- 'bash tools/round2_two_gpu.sh'):
# the policy step with the gradient exchange after backward (default) vs started from inside backward
# (V2A_OVERLAP_ALLREDUCE=1), same box, interleaved; then the default bench line at N = 2.
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for rep in 1 2; do
  for ov in 0 1; do
    V2A_OVERLAP_ALLREDUCE=$ov timeout 120 $RUN --master-port $((29500 + rep * 2 + ov)) tools/online_loop.py \
        --tasks 2 --policy-steps 40 > gpurun_out/r2_overlap_${ov}_${rep}.json 2> gpurun_out/r2_overlap_${ov}_${rep}.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_overlap_${ov}_${rep}.json").read().strip().splitlines()[-1])
    print("overlap=$ov rep=$rep train_ms_per_step", round(d["train_ms_per_step"], 3), "loss", d["loss"])
except Exception as e:
    print("overlap=$ov rep=$rep failed", e)
PY
  done
done
timeout 280 $RUN --master-port 29600 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
cut -c1-300 gpurun_out/r2_bench_n2.json
