#!/bin/bash
# A/B: second L2 prefetch (the middle line of the child records) when a level is read in three batches
mkdir -p gpurun_out
prev=$PWD/synthesis_b200/libsynthesis_b200_prev.so
q="--steps 1 --warmup 1 --games-mult 2 --no-cpu-baseline"
for v in prev new prev new; do
  if [ $v = prev ]; then export SYN_B200_LIB=$prev; else unset SYN_B200_LIB; fi
  timeout 200 python bench.py $q > gpurun_out/nn_pf_$v.json 2> gpurun_out/nn_pf_$v.err
  echo "nn $v: $(python -c "import json;d=json.load(open('gpurun_out/nn_pf_$v.json'));print(round(d['value']/1e6,1))")"
done
