#!/bin/bash
# Round-2 eight-GPU call (SURVEY configs[3] and [4]): the default bench line under torchrun on 8 ranks -- 16 videos per
# rank (no collective), the policy optimisation step at B = 256 per rank with the gradient all-reduce started from
# inside backward (NCCL AVG over NVSwitch), and the online loop with one task per rank.
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 1100 $RUN --master-port 29531 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r2n8_bench.json 2> gpurun_out/r2n8_bench.err; echo "bench rc=$?"
tail -c 1400 gpurun_out/r2n8_bench.json; grep -v "sampling loop" gpurun_out/r2n8_bench.err | tail -5
