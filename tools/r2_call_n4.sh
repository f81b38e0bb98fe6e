#!/bin/bash
# Round-2 four-GPU line of the final tree (quick: no extra legs).
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29551 bench.py --gpus 4 --steps 2 --warmup 3 --no-extras > gpurun_out/r2n4_bench.json 2> gpurun_out/r2n4_bench.err; echo "bench rc=$?"
tail -c 500 gpurun_out/r2n4_bench.json
