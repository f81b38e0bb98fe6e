#!/bin/bash
# Round-2 GPU call 26: final tree -- smoke(), default bench line, reference arm as the driver runs it.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c26_smoke.log 2>&1; tail -2 gpurun_out/r2c26_smoke.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2c26_bench.json 2> gpurun_out/r2c26_bench.err; echo "bench rc=$?"
tail -c 1300 gpurun_out/r2c26_bench.json
