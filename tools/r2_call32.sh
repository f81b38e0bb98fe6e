#!/bin/bash
# Round-2 GPU call 32: last full validation of the round's final tree: GPU suite + smoke().
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c32_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c32_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c32_smoke.log 2>&1; tail -2 gpurun_out/r2c32_smoke.log
