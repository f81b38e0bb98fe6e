#!/bin/bash
# Round-2 GPU call 36: encoder / policy suites after the one-pass weight-gradient stage fix.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py tests/test_policy_gpu.py -m gpu -q > gpurun_out/r2c36_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c36_pytest.log
