# after taking the L2-priority plumbing out again: GPU suite, bench arm, phase clocks
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
( time timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; cut -c1-260 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; tail -1 gpurun_out/launches_bench.log | cut -c1-200
SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 454656 800 1 nn | tee gpurun_out/phases_bench.log
