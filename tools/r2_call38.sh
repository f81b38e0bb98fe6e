#!/bin/bash
# Round-2 GPU call 38: the round's last tree once more: GPU suite, smoke(), default bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c38_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c38_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c38_smoke.log 2>&1; tail -2 gpurun_out/r2c38_smoke.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2c38_bench.json 2> gpurun_out/r2c38_bench.err; echo "bench rc=$?"
tail -c 900 gpurun_out/r2c38_bench.json
