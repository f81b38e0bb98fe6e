#!/bin/bash
# Round-2 two-GPU call: the default bench line under torchrun (policy all-reduce started from inside backward, NCCL AVG;
# online loop on 2 ranks), then the same policy step with the blocking exchange for the A/B.
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $RUN --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2n2_bench.json 2> gpurun_out/r2n2_bench.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/r2n2_bench.json; grep -v "sampling loop" gpurun_out/r2n2_bench.err | tail -5
for ov in 1 0 1 0; do
  V2A_OVERLAP_ALLREDUCE=$ov timeout 200 $RUN --master-port $((29520 + RANDOM % 50)) tools/online_loop.py --tasks 2 --policy-steps 40 > gpurun_out/r2n2_overlap_${ov}.json 2> gpurun_out/r2n2_overlap_${ov}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2n2_overlap_${ov}.json").read().strip().splitlines()[-1])
    print("overlap=$ov train_ms_per_step", round(d["train_ms_per_step"], 3), "loss", d["loss"])
except Exception as e:
    print("overlap=$ov failed", e)
PY
done
