#!/bin/bash
# Round-2 GPU call 1 (one GPU): full GPU test suite on the new tree, per-layer timings with the sub-pixel upsample
# convs, the default bench line (with the gpu_eager / fast / online_loop legs).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2c1_gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c1_pytest.log
tail -5 gpurun_out/r2c1_pytest.log
timeout 200 python tools/quick_bench.py 16 --layers > gpurun_out/r2c1_layers.txt 2>&1; echo "layers rc=$?"
head -4 gpurun_out/r2c1_layers.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?"
tail -c 1600 gpurun_out/r2c1_bench.json
tail -5 gpurun_out/r2c1_bench.err
