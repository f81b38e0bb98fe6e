#!/bin/bash
# Round-2 GPU call 11: tensor-core (mma.sync) attention kernel: kernel test, network goldens, A/B.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "attention" > gpurun_out/r2c11_att.log 2>&1; echo "attention rc=$?"; tail -5 gpurun_out/r2c11_att.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c11_pytest.log 2>&1; echo "pytest rc=$?"
grep -v "sampling loop" gpurun_out/r2c11_pytest.log | grep -E "passed|failed|^FAILED|^E  " | tail -12
timeout 600 python tools/ab_forward.py V2A_ATTN_MMA=0 V2A_ATTN_MMA=1 > gpurun_out/r2c11_ab.txt 2>&1; cat gpurun_out/r2c11_ab.txt
timeout 200 python tools/quick_bench.py 16 --layers > gpurun_out/r2c11_layers.txt 2>&1; grep "by kind" -A 12 gpurun_out/r2c11_layers.txt
