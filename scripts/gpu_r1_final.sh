# round-1 evidence run at HEAD (one B200): GPU tests, both bench arms, ncu launch list of the bench command, DRAM traffic of
# one bench launch, phase clocks, and the measurements of the widened rows (deduplicate, learner, evaluation sweep)
mkdir -p gpurun_out
nproc; free -g | head -2 | tail -1
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
( time timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; cut -c1-300 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
( time timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2>&1 | grep real; cut -c1-200 gpurun_out/bench_reference.json
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; tail -1 gpurun_out/launches_bench.log | cut -c1-200
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:selfplay -c 1 --csv --log-file gpurun_out/traffic_nn_e800_g454656.csv python scripts/prof_driver.py 454656 800 1 nn > gpurun_out/traffic.log 2>&1; tail -2 gpurun_out/traffic.log
SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 454656 800 1 nn | tee gpurun_out/phases_bench.log
timeout 600 python scripts/bench_dedup.py 454656 64 5 2>&1 | tail -1 | tee gpurun_out/dedup_454k.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:insert_kernel|rep_kernel|rs_|scan_|heads_|starts_|reduce_|fill_kernel|gather_kernel" --csv --log-file gpurun_out/launches_dedup.csv python scripts/bench_dedup.py 454656 64 1 > gpurun_out/launches_dedup.log 2>&1
SYN_TRAIN_PROF=1 timeout 600 python scripts/bench_train.py 1000 200 20000 2000 2>&1 | grep -E "train phases|op" | tail -2 | tee gpurun_out/train_bench.log
