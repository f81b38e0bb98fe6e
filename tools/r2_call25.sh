#!/bin/bash
# Round-2 GPU call 25: programmatic dependent launch on the tensor-core kernels (igemm, dual, wgrad): full GPU suite,
# A/B of the video forward and of the policy optimisation step.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c25_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c25_pytest.log
timeout 600 python tools/ab_forward.py V2A_PDL=1 V2A_PDL=0 > gpurun_out/r2c25_ab.txt 2>&1; cat gpurun_out/r2c25_ab.txt
for v in 1 0 1 0; do
  echo "V2A_PDL=$v"; V2A_PDL=$v timeout 200 python tools/quick_bench_loss.py > gpurun_out/r2c25_pol_$v.txt 2>&1; tail -1 gpurun_out/r2c25_pol_$v.txt
done
