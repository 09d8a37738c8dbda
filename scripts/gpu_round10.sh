#!/bin/bash
# final check at HEAD: full GPU suite, smoke, evaluation-sweep throughput (match kernel with batched child loads)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_head.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests_head.log; tail -3 gpurun_out/gpu_tests_head.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 python scripts/bench_match.py > gpurun_out/match_bench.txt 2> gpurun_out/match_bench.err; cut -c1-400 gpurun_out/match_bench.txt
