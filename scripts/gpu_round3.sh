#!/bin/bash
# NN kernel: teams 4 / 5 / 6 at the default bench configuration; ncu launch list + one full capture of the rollout kernel
mkdir -p gpurun_out
for t in 4 5 6; do
  SYN_TPG_TEAMS=$t timeout 400 python bench.py --no-cpu-baseline > gpurun_out/nn_default_teams_$t.json 2> gpurun_out/nn_default_teams_$t.err
  echo "nn teams $t (default bench): $(python -c "import json;d=json.load(open('gpurun_out/nn_default_teams_$t.json'));print(round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1))")"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_rollout_bench.csv python bench.py --leaf rollout --steps 1 --warmup 1 --games-mult 1 --no-cpu-baseline > gpurun_out/ncu_rollout_bench.log 2>&1
echo "ncu launch list exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:selfplay_rollout_tpg2 -c 1 -f -o gpurun_out/rollout_tpg2_full python scripts/prof_driver.py 151552 100 1 rollout > gpurun_out/ncu_rollout_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
