#!/bin/bash
# One GPU call: parity tests that exercise the rollout thread-per-game kernel, then configs[0] throughput per variant.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rollout or survey or group_lanes or fixtures or stop_games or reductions or threads_per_cta or sharding or error" > gpurun_out/rollout_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/rollout_tests.log
tail -5 gpurun_out/rollout_tests.log
for t in 512 640 768 1024; do
  SYN_ROLLOUT_THREADS=$t timeout 300 python bench.py --leaf rollout --steps 1 --warmup 1 --games-mult 2 --no-cpu-baseline > gpurun_out/rollout_bench_$t.json 2> gpurun_out/rollout_bench_$t.err
  echo "threads $t exit $?"; cut -c1-200 gpurun_out/rollout_bench_$t.json
done
SYN_GROUP_LANES=16 timeout 300 python bench.py --leaf rollout --steps 1 --warmup 1 --games-mult 2 --no-cpu-baseline > gpurun_out/rollout_bench_gl16.json 2> gpurun_out/rollout_bench_gl16.err
cut -c1-200 gpurun_out/rollout_bench_gl16.json
