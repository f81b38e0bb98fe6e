#!/bin/bash
# Round-2 GPU call 9: the default bench line (all legs) on the current tree + the reference arm as the driver runs it.
mkdir -p gpurun_out
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2c9_ref.json 2> gpurun_out/r2c9_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2c9_ref.json
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2c9_bench.json; grep -v "sampling loop" gpurun_out/r2c9_bench.err | tail -3
