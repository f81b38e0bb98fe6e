#!/bin/bash
# Round-2 GPU call 29: B = 1 video forward (the per-rank shape of configs[4] on 8 GPUs) with programmatic dependent
# launch on / off; policy tests with the side-lane residual conv as default.
mkdir -p gpurun_out
for v in 1 0 1 0; do
  echo "V2A_PDL=$v"; V2A_PDL=$v timeout 200 python tools/quick_bench.py 1 > gpurun_out/r2c29_b1_$v.txt 2>&1; sed -n 2,2p gpurun_out/r2c29_b1_$v.txt
done
timeout 600 python -m pytest tests/test_policy_gpu.py tests/test_encoder_gpu.py -m gpu -q > gpurun_out/r2c29_pytest.log 2>&1; tail -2 gpurun_out/r2c29_pytest.log
