#!/bin/bash
# Round-2 GPU call 10: linear-tile epilogue set-up: GPU suite + A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c10_pytest.log 2>&1; echo "pytest rc=$?"
grep -v "sampling loop" gpurun_out/r2c10_pytest.log | grep -E "passed|failed|^FAILED|^E  " | tail -12
timeout 600 python tools/ab_forward.py V2A_LINEAR=0 V2A_LINEAR=1 > gpurun_out/r2c10_ab.txt 2>&1; cat gpurun_out/r2c10_ab.txt
timeout 200 python tools/quick_bench.py 16 --layers > gpurun_out/r2c10_layers.txt 2>&1; sed -n 2,18p gpurun_out/r2c10_layers.txt
timeout 200 python tools/quick_bench_policy.py 256 > gpurun_out/r2c10_policy.txt 2>&1; head -4 gpurun_out/r2c10_policy.txt
timeout 200 python tools/quick_bench_encoder.py 256 > gpurun_out/r2c10_encoder.txt 2>&1; head -3 gpurun_out/r2c10_encoder.txt
