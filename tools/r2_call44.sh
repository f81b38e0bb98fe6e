#!/bin/bash
# Round-2 GPU call 44: split-K slices + deterministic reduction: ops / video suites, smoke.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_video_gpu.py -m gpu -q > gpurun_out/r2c44_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "sampling loop" gpurun_out/r2c44_pytest.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c44_smoke.log 2>&1; tail -2 gpurun_out/r2c44_smoke.log
