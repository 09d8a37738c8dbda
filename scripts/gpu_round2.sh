#!/bin/bash
# full GPU suite at HEAD, rollout variants, NN team variants, default rollout bench with CPU baseline
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_head.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests_head.log; tail -3 gpurun_out/gpu_tests_head.log
q="--steps 1 --warmup 1 --games-mult 2 --no-cpu-baseline"
for v in "1024 3" "1024 5" "896 3" "896 5"; do set -- $v
  SYN_ROLLOUT_THREADS=$1 SYN_ROLLOUT_CW=$2 timeout 300 python bench.py --leaf rollout $q > gpurun_out/rb_$1_$2.json 2> gpurun_out/rb_$1_$2.err
  echo "rollout threads $1 cw $2: $(python -c "import json;d=json.load(open('gpurun_out/rb_$1_$2.json'));print(round(d['value']/1e6,1))")"
done
SYN_TPG_FAST_SELECT=1 timeout 300 python bench.py --leaf rollout $q > gpurun_out/rb_fast.json 2> gpurun_out/rb_fast.err
echo "rollout 1024/3 fast select: $(python -c "import json;d=json.load(open('gpurun_out/rb_fast.json'));print(round(d['value']/1e6,1))")"
for f in 0 1; do
  SYN_TPG_FAST_SELECT=$f timeout 300 python bench.py $q > gpurun_out/nn_fast_$f.json 2> gpurun_out/nn_fast_$f.err
  echo "nn teams 4 fast select $f: $(python -c "import json;d=json.load(open('gpurun_out/nn_fast_$f.json'));print(round(d['value']/1e6,1))")"
done
for t in 5 8; do
  SYN_TPG_TEAMS=$t timeout 300 python bench.py $q > gpurun_out/nn_teams_$t.json 2> gpurun_out/nn_teams_$t.err
  echo "nn teams $t: $(python -c "import json;d=json.load(open('gpurun_out/nn_teams_$t.json'));print(round(d['value']/1e6,1))")"
done
timeout 600 python bench.py --leaf rollout > gpurun_out/bench_rollout_n1.json 2> gpurun_out/bench_rollout_n1.err; cut -c1-150 gpurun_out/bench_rollout_n1.json
