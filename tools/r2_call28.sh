#!/bin/bash
# Round-2 GPU call 28: residual 1x1 conv on the side lane also at training batch (V2A_SIDE_RES=1): UNet1D step, loss step.
mkdir -p gpurun_out
for v in 0 1 0 1; do
  echo "V2A_SIDE_RES=$v"
  V2A_SIDE_RES=$v timeout 200 python tools/quick_bench_policy.py 256 > gpurun_out/r2c28_unet_$v.txt 2>&1; sed -n 2,3p gpurun_out/r2c28_unet_$v.txt
  V2A_SIDE_RES=$v timeout 200 python tools/quick_bench_loss.py > gpurun_out/r2c28_loss_$v.txt 2>&1; tail -1 gpurun_out/r2c28_loss_$v.txt
done
V2A_SIDE_RES=1 timeout 600 python -m pytest tests/test_policy_gpu.py -m gpu -q > gpurun_out/r2c28_pytest.log 2>&1; tail -2 gpurun_out/r2c28_pytest.log
