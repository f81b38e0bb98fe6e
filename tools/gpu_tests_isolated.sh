#!/bin/bash
# Run each GPU test function in its own process so one kernel trap cannot poison the rest.
# usage: tools/gpu_tests_isolated.sh tests/test_ops_gpu.py [more files]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
for f in "$@"; do
  for t in $(python -m pytest "$f" --collect-only -q -m gpu 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u); do
    name=$(echo "$t" | tr '/:' '__')
    timeout 300 python -m pytest "$t" -q -m gpu -x --timeout 240 > "gpurun_out/$name.log" 2>&1
    rc=$?
    echo "rc=$rc $t :: $(tail -1 gpurun_out/$name.log)"
    if [ $rc -ne 0 ]; then grep -E "^E |v2a:|Error|error" "gpurun_out/$name.log" | head -12; fi
  done
done
