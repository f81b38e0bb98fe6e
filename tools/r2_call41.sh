#!/bin/bash
# Round-2 GPU call 41: small-grid rule tuning at B = 1 / 2 (fill target, N-tile floor).
mkdir -p gpurun_out
for s in "0.7 64" "0.5 64" "0.9 64" "0.7 32" "0.9 32" "1.2 64" "0.7 64"; do
  set -- $s
  echo "fill=$1 floor=$2: $(V2A_BN_FILL=$1 V2A_BN_FLOOR=$2 timeout 200 python tools/quick_bench.py 1 2>&1 | sed -n 2,2p | cut -c1-40)  |  B=2: $(V2A_BN_FILL=$1 V2A_BN_FLOOR=$2 timeout 200 python tools/quick_bench.py 2 2>&1 | sed -n 2,2p | cut -c1-40)"
done | tee gpurun_out/r2c41_bn.txt
