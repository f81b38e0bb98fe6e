#!/bin/bash
# Round-2 GPU call 8: compile-time specialised epilogues (kEpi): full GPU suite, A/B of the B = 16 forward
# (V2A_FAST_EPILOGUE=0/1 x V2A_DUAL=0/1), per-layer tables, predict_action.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c8_pytest.log 2>&1; echo "pytest rc=$?"
grep -v "sampling loop" gpurun_out/r2c8_pytest.log | grep -E "passed|failed|^FAILED|^E  " | tail -12
timeout 600 python tools/ab_forward.py V2A_FAST_EPILOGUE=0,V2A_DUAL=0 V2A_FAST_EPILOGUE=1,V2A_DUAL=0 V2A_FAST_EPILOGUE=1,V2A_DUAL=1 > gpurun_out/r2c8_ab.txt 2>&1; cat gpurun_out/r2c8_ab.txt
V2A_DUAL=0 timeout 200 python tools/quick_bench.py 16 --layers > gpurun_out/r2c8_layers_nodual.txt 2>&1; sed -n 2,18p gpurun_out/r2c8_layers_nodual.txt; grep "cout   128" gpurun_out/r2c8_layers_nodual.txt | grep "#"
V2A_DUAL=1 timeout 200 python tools/quick_bench.py 16 --layers > gpurun_out/r2c8_layers_dual.txt 2>&1; sed -n 2,4p gpurun_out/r2c8_layers_dual.txt; grep "cout   128" gpurun_out/r2c8_layers_dual.txt | grep "#"
timeout 120 python tools/quick_bench_predict.py > gpurun_out/r2c8_predict.txt 2>&1; grep -v "UNet1D forward graph" gpurun_out/r2c8_predict.txt | tail -8
timeout 200 python tools/quick_bench_policy.py 256 > gpurun_out/r2c8_policy.txt 2>&1; head -4 gpurun_out/r2c8_policy.txt
