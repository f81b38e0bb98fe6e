#!/bin/bash
# Round-2 GPU call 37: max-pool backward evaluated inside the stem's GroupNorm backward: parity, step time A/B.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py tests/test_policy_gpu.py -m gpu -q > gpurun_out/r2c37_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c37_pytest.log
for v in 1 0 1 0; do
  echo "V2A_ENC_FUSE_POOL=$v"; V2A_ENC_FUSE_POOL=$v timeout 200 python tools/quick_bench_loss.py > gpurun_out/r2c37_loss_$v.txt 2>&1; tail -1 gpurun_out/r2c37_loss_$v.txt
done
V2A_ENC_FUSE_POOL=1 timeout 200 python tools/quick_bench_encoder.py 256 --layers > gpurun_out/r2c37_enc.txt 2>&1; sed -n 2,2p gpurun_out/r2c37_enc.txt; grep "by kind" gpurun_out/r2c37_enc.txt | tail -1
