#!/bin/bash
# rollout kernel with the ChaCha key kept in shared memory: parity, then A/B against the previous build
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "rollout or survey or group_lanes or fixtures or stop_games or reductions or threads_per_cta or sharding or config3" > gpurun_out/gpu_tests_keys.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests_keys.log; tail -3 gpurun_out/gpu_tests_keys.log
q="--leaf rollout --steps 1 --warmup 1 --games-mult 3 --no-cpu-baseline"
prev=$PWD/synthesis_b200/libsynthesis_b200_prev.so
for v in prev keys prev keys; do
  if [ $v = prev ]; then export SYN_B200_LIB=$prev; else unset SYN_B200_LIB; fi
  timeout 300 python bench.py $q > gpurun_out/rb_$v.json 2> gpurun_out/rb_$v.err
  echo "rollout $v: $(python -c "import json;d=json.load(open('gpurun_out/rb_$v.json'));print(round(d['value']/1e6,1))")"
done
