#!/bin/bash
# parity of the counter / won() changes, then A/B: mbarrier back-off (two extra builds of the same source), per-warp counters
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_r5.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests_r5.log; tail -3 gpurun_out/gpu_tests_r5.log
q="--steps 1 --warmup 1 --games-mult 3 --no-cpu-baseline"
for v in "" _sleep40 _sleep150; do
  lib=$PWD/synthesis_b200/libsynthesis_b200$v.so
  SYN_B200_LIB=$lib timeout 300 python bench.py $q > gpurun_out/nn_mbar$v.json 2> gpurun_out/nn_mbar$v.err
  echo "nn lib '$v': $(python -c "import json;d=json.load(open('gpurun_out/nn_mbar$v.json'));print(round(d['value']/1e6,1))")"
done
timeout 300 python bench.py --leaf rollout $q > gpurun_out/rb_r5.json 2> gpurun_out/rb_r5.err
echo "rollout: $(python -c "import json;d=json.load(open('gpurun_out/rb_r5.json'));print(round(d['value']/1e6,1))")"
