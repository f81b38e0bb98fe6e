#!/bin/bash
# Round-2 GPU call 43: deep spatial convs at small batch as split-K into an fp32 scratch + plane split: parity, timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_video_gpu.py -m gpu -q > gpurun_out/r2c43_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "sampling loop" gpurun_out/r2c43_pytest.log | tail -3
for v in 1 0 1 0; do
  echo "V2A_SPLITK_SPATIAL=$v: $(V2A_SPLITK_SPATIAL=$v timeout 200 python tools/quick_bench.py 1 2>&1 | sed -n 2,2p | cut -c1-24) | $(V2A_SPLITK_SPATIAL=$v timeout 200 python tools/quick_bench.py 2 2>&1 | sed -n 2,2p | cut -c1-24) | $(V2A_SPLITK_SPATIAL=$v timeout 200 python tools/quick_bench.py 4 2>&1 | sed -n 2,2p | cut -c1-24)"
done | tee gpurun_out/r2c43_ab.txt
