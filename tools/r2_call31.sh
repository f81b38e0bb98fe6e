#!/bin/bash
# Round-2 GPU call 31: small-grid rule for the N tile (video engine at small batch): parity tests, B = 1 / 2 / 16 forward.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_video_gpu.py tests/test_ops_gpu.py -m gpu -q > gpurun_out/r2c31_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "sampling loop" gpurun_out/r2c31_pytest.log | tail -3
for b in 1 2 4 16; do timeout 200 python tools/quick_bench.py $b > gpurun_out/r2c31_b$b.txt 2>&1; sed -n 2,2p gpurun_out/r2c31_b$b.txt; done
timeout 200 python tools/quick_bench.py 1 --layers > gpurun_out/r2c31_layers_b1.txt 2>&1; sed -n 5,12p gpurun_out/r2c31_layers_b1.txt
