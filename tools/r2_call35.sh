#!/bin/bash
# Round-2 GPU call 35: weight-gradient GEMMs with one bf16 product (long reductions): gradient movement at B = 256, step time.
mkdir -p gpurun_out
timeout 900 python tools/wgrad_passes_probe.py 1 auto auto:1024 > gpurun_out/r2c35_probe.txt 2>&1; cat gpurun_out/r2c35_probe.txt
for v in 3 1 auto 3 1; do
  echo "V2A_WGRAD_PASSES=$v"; V2A_WGRAD_PASSES=$v timeout 200 python tools/quick_bench_loss.py > gpurun_out/r2c35_loss_$v.txt 2>&1; tail -1 gpurun_out/r2c35_loss_$v.txt
done
