#!/bin/bash
# rollout kernel, 32-word ring at every thread count: does the larger L1 change the ranking of the variants?
mkdir -p gpurun_out
q="--leaf rollout --steps 1 --warmup 1 --games-mult 3 --no-cpu-baseline"
for t in 1024 768 896 640 512 1024; do
  SYN_ROLLOUT_THREADS=$t timeout 300 python bench.py $q > gpurun_out/rb32_$t.json 2> gpurun_out/rb32_$t.err
  echo "rollout threads $t ring 32: $(python -c "import json;d=json.load(open('gpurun_out/rb32_$t.json'));print(round(d['value']/1e6,1))")"
done
