#!/bin/bash
# last check at HEAD: full GPU suite + smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_head.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests_head.log; tail -3 gpurun_out/gpu_tests_head.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
