#!/bin/bash
# Round-2 GPU call 45: full GPU suite and a quick bench line on the last tree (split-K slices for small batches added).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c45_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c45_pytest.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2c45_bench.json 2> gpurun_out/r2c45_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2c45_bench.json
