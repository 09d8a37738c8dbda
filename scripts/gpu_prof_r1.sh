# Round-1 profile artefacts (numbers printed under ncu are never bench values)
mkdir -p gpurun_out
# (1) launch list of the bench command on a reduced step size
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 2 --warmup 1 --games-mult 1 --no-cpu-baseline > gpurun_out/r1_launches_bench.log 2>&1; tail -2 gpurun_out/r1_launches_bench.log | cut -c1-300
# (2) one full capture of the dominant kernel, all SMs busy, short trees (E=100)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:selfplay -c 1 -o gpurun_out/r1_nn_tpg8_full python scripts/prof_driver.py 151552 100 1 nn > gpurun_out/r1_nn_tpg8_full.log 2>&1; tail -3 gpurun_out/r1_nn_tpg8_full.log
