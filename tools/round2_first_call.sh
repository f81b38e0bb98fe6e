#!/bin/bash
# First GPU call of the next round (one GPU, ~6 min): the evidence round 1 ran out of budget for.
#   gpurun --timeout 420 -- 'bash tools/round2_first_call.sh'
# 1. full GPU test suite on the final tree
# 2. default bench line (now carries policy.e2e_device_replay)
# 3. ncu --set full of the kernels added at the end of round 1 (replay gather, PerceiverResampler pieces)
# 4. the 1-pass numerics class (V2A_PASSES=1) next to the default, same box: UNet forward time at B=16
# 5. online loop with the videos of all tasks batched in one sample() call
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest.log
timeout 240 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
timeout 90 $NCU --set full --import-source on -k regex:"replay_gather" -c 4 -f -o gpurun_out/full_replay \
    python tools/quick_bench_replay.py 256 > gpurun_out/full_replay.log 2>&1
STEPS=1 B=16 timeout 120 $NCU --set full --import-source on -k regex:"pr_" -c 40 -f -o gpurun_out/full_perceiver \
    python tools/profile_target.py > gpurun_out/full_perceiver.log 2>&1
timeout 120 python tools/ab_forward.py V2A_PASSES=3 V2A_PASSES=1 > gpurun_out/r2_passes_ab.txt 2>&1 || true
timeout 120 python tools/online_loop.py --tasks 8 --policy-steps 10 --batch-videos > gpurun_out/r2_online_batched.json 2> gpurun_out/r2_online_batched.err
tail -3 gpurun_out/r2_pytest.log; cut -c1-300 gpurun_out/r2_bench.json; cat gpurun_out/r2_passes_ab.txt | tail -5; cat gpurun_out/r2_online_batched.json
du -sh gpurun_out
