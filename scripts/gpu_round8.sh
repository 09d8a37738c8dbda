#!/bin/bash
# steady-state lane utilisation / issue / L2 metrics of ONE launch of each bench configuration (few-pass ncu metrics, not --set full)
mkdir -p gpurun_out
M=smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
timeout 400 ncu --metrics $M --clock-control none -k regex:selfplay -c 1 --csv --log-file gpurun_out/steady_nn_e800_g568320.csv python scripts/prof_driver.py 568320 800 1 nn > gpurun_out/steady_nn.log 2>&1; echo "nn exit $?"
timeout 400 ncu --metrics $M --clock-control none -k regex:selfplay -c 1 --csv --log-file gpurun_out/steady_rollout_e800_g909312.csv python scripts/prof_driver.py 909312 800 1 rollout > gpurun_out/steady_rollout.log 2>&1; echo "rollout exit $?"
grep -h "selfplay" gpurun_out/steady_*.csv | cut -d, -f5,13-15
