#!/bin/bash
# evidence at HEAD: full GPU suite, smoke, default bench, rollout bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_head.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests_head.log; tail -3 gpurun_out/gpu_tests_head.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_default_n1.json 2> gpurun_out/bench_default_n1.err; cut -c1-160 gpurun_out/bench_default_n1.json
timeout 600 python bench.py --leaf rollout > gpurun_out/bench_rollout_n1.json 2> gpurun_out/bench_rollout_n1.err; cut -c1-160 gpurun_out/bench_rollout_n1.json
