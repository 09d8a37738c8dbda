# round 2, GPU call 17: the record runs at HEAD (8b7cab9): bench.py on every config (both arms), DRAM traffic captures, launch list
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err; tail -c 600 gpurun_out/r2_bench_default_n1.json
for c in 0 2 4; do
  timeout 900 python bench.py --config $c > gpurun_out/r2_bench_config${c}_n1.json 2> gpurun_out/r2_bench_config${c}_n1.err; head -c 300 gpurun_out/r2_bench_config${c}_n1.json; echo
done
timeout 900 python bench.py --leaf rollout --steps 1 --warmup 1 > gpurun_out/r2_bench_rollout_n1.json 2> gpurun_out/r2_bench_rollout_n1.err; head -c 300 gpurun_out/r2_bench_rollout_n1.json; echo
for c in 1 0 2 4; do
  timeout 900 python bench.py --impl reference --config $c --steps 1 --warmup 1 > gpurun_out/r2_bench_reference_config${c}.json 2> gpurun_out/r2_bench_reference_config${c}.err; head -c 260 gpurun_out/r2_bench_reference_config${c}.json; echo
done
# DRAM traffic of ONE launch of the default bench configuration and of configs[2] (for roofline.traffic)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:selfplay -c 1 --csv --log-file gpurun_out/r2_traffic_nn_e800_g568320.csv python scripts/prof_driver.py 568320 800 0 nn > gpurun_out/r2_traffic_nn.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:selfplay -c 1 --csv --log-file gpurun_out/r2_traffic_nn_e1600_g4096.csv python scripts/prof_driver.py 4096 1600 0 nn 4096 > gpurun_out/r2_traffic_c2.log 2>&1
# launch list of the bench command (cold, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_n1.csv python bench.py --steps 1 --warmup 1 --games-mult 1 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
tail -3 gpurun_out/r2_traffic_nn.log gpurun_out/r2_traffic_c2.log | cut -c1-200
