#!/bin/bash
# Round-2 GPU call 3: full GPU suite again (CFG fused path, .data cache tests), quick bench without the extras.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2c3_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c3_pytest.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2c3_pytest.log | tail -20
timeout 400 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err; echo "bench rc=$?"
tail -c 900 gpurun_out/r2c3_bench.json; tail -3 gpurun_out/r2c3_bench.err
