#!/bin/bash
# Evidence at HEAD: full GPU suite, default bench (NN, 5 teams), reference arm, rollout bench, ncu launch list,
# DRAM traffic of one launch of each bench configuration, one full capture of the NN kernel, phase clocks.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_head.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests_head.log; tail -3 gpurun_out/gpu_tests_head.log
timeout 600 python bench.py > gpurun_out/bench_default_n1.json 2> gpurun_out/bench_default_n1.err; cut -c1-160 gpurun_out/bench_default_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference_n1.json 2> gpurun_out/bench_reference_n1.err; cut -c1-160 gpurun_out/bench_reference_n1.json
timeout 600 python bench.py --leaf rollout > gpurun_out/bench_rollout_n1.json 2> gpurun_out/bench_rollout_n1.err; cut -c1-160 gpurun_out/bench_rollout_n1.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_bench_n1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "launch list exit $?"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none -k regex:selfplay -c 1 --csv --log-file gpurun_out/traffic_nn_e800_g568320.csv python scripts/prof_driver.py 568320 800 1 nn > gpurun_out/traffic_nn.log 2>&1; echo "traffic nn exit $?"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none -k regex:selfplay -c 1 --csv --log-file gpurun_out/traffic_rollout_e800_g909312.csv python scripts/prof_driver.py 909312 800 1 rollout > gpurun_out/traffic_rollout.log 2>&1; echo "traffic rollout exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:selfplay_nn_tpg2 -c 1 -f -o gpurun_out/nn_tpg2_t5_full python scripts/prof_driver.py 94720 200 1 nn > gpurun_out/ncu_nn_full.log 2>&1; echo "ncu full exit $?"
SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 568320 800 1 nn > gpurun_out/phases_bench.txt 2>&1; cat gpurun_out/phases_bench.txt
