#!/bin/bash
# Round-2 GPU call 42: the reference arm as the driver runs it, on the final tree.
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c42_ref.json 2> gpurun_out/r2c42_ref.err; echo "ref rc=$?"
tail -c 1200 gpurun_out/r2c42_ref.json
