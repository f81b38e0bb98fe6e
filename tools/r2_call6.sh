#!/bin/bash
# Round-2 GPU call 6: pipelined TMEM reads in the fused-split epilogues, small-M rewrite: full GPU suite, dual A/B,
# per-layer table, predict_action breakdown.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c6_pytest.log 2>&1; echo "pytest rc=$?"
grep -v "sampling loop" gpurun_out/r2c6_pytest.log | grep -E "passed|failed|^FAILED|^E  " | tail -12
timeout 400 python tools/ab_forward.py V2A_DUAL=0 V2A_DUAL=1 > gpurun_out/r2c6_ab.txt 2>&1; cat gpurun_out/r2c6_ab.txt
timeout 200 python tools/quick_bench.py 16 --layers > gpurun_out/r2c6_layers.txt 2>&1; sed -n 2,4p gpurun_out/r2c6_layers.txt; grep "cout   128" gpurun_out/r2c6_layers.txt | grep "#"
V2A_DUAL=0 timeout 200 python tools/quick_bench.py 16 --layers > gpurun_out/r2c6_layers_nodual.txt 2>&1; sed -n 2,4p gpurun_out/r2c6_layers_nodual.txt; grep "cout   128" gpurun_out/r2c6_layers_nodual.txt | grep "#"
timeout 120 python tools/quick_bench_predict.py > gpurun_out/r2c6_predict.txt 2>&1; cat gpurun_out/r2c6_predict.txt | tail -10
