#!/bin/bash
# Round-2 two-GPU re-check of the final tree (programmatic dependent launches + NCCL in one process).
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $RUN --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2n2b_bench.json 2> gpurun_out/r2n2b_bench.err; echo "bench rc=$?"
tail -c 900 gpurun_out/r2n2b_bench.json; grep -v "sampling loop" gpurun_out/r2n2b_bench.err | tail -3
