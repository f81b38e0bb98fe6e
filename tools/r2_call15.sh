#!/bin/bash
# Round-2 GPU call 15: one-pass cluster GroupNorm backward of the encoders: parity tests, A/B against the two-pass form.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py tests/test_policy_gpu.py -m gpu -q > gpurun_out/r2c15_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c15_pytest.log
for v in 1 0 1 0; do
  echo "V2A_ENC_GN_CLUSTER=$v"
  V2A_ENC_GN_CLUSTER=$v timeout 200 python tools/quick_bench_encoder.py 256 --layers > gpurun_out/r2c15_enc_$v.txt 2>&1; sed -n 2,2p gpurun_out/r2c15_enc_$v.txt; grep "by kind" gpurun_out/r2c15_enc_$v.txt | tail -1
  V2A_ENC_GN_CLUSTER=$v timeout 200 python tools/quick_bench_loss.py > gpurun_out/r2c15_pol_$v.txt 2>&1; tail -1 gpurun_out/r2c15_pol_$v.txt
done
