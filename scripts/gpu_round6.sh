#!/bin/bash
# A/B: two TMEM reads in flight per wait in the forward's epilogue (-DSYN_TMEM_PAIR=1 build of the same source)
mkdir -p gpurun_out
pair=$PWD/synthesis_b200/libsynthesis_b200_pair.so
SYN_B200_LIB=$pair timeout 900 python -m pytest tests -x -q -m gpu -k "nn or match or teams or eval or train or alpha" > gpurun_out/gpu_tests_pair.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests_pair.log; tail -3 gpurun_out/gpu_tests_pair.log
q="--steps 1 --warmup 1 --games-mult 3 --no-cpu-baseline"
for v in base pair base pair; do
  if [ $v = pair ]; then export SYN_B200_LIB=$pair; else unset SYN_B200_LIB; fi
  timeout 300 python bench.py $q > gpurun_out/nn_$v.json 2> gpurun_out/nn_$v.err
  echo "nn $v: $(python -c "import json;d=json.load(open('gpurun_out/nn_$v.json'));print(round(d['value']/1e6,1))")"
done
