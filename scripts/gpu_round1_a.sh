# GPU pass A: parity tests, bench (both leaf modes + reference arm), ncu launch list + one full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; lscpu | grep -E "Model name|^CPU\(s\)"
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_nn.json 2> gpurun_out/bench_nn.err; cat gpurun_out/bench_nn.json; tail -3 gpurun_out/bench_nn.err
timeout 600 python bench.py --steps 3 --warmup 3 --leaf rollout --no-cpu-baseline > gpurun_out/bench_ro.json 2> gpurun_out/bench_ro.err; cat gpurun_out/bench_ro.json; tail -3 gpurun_out/bench_ro.err
timeout 600 python bench.py --steps 3 --warmup 3 --group-lanes 32 --no-cpu-baseline > gpurun_out/bench_nn_gl32.json 2> gpurun_out/bench_nn_gl32.err; cat gpurun_out/bench_nn_gl32.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>&1; cat gpurun_out/bench_reference.json
# launch list of the bench command (numbers under ncu are never bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --games 1184 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; tail -2 gpurun_out/launches_bench.log
# one full capture of the dominant kernel (short workload)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:selfplay -c 1 -o gpurun_out/prof_r1_nn python scripts/prof_driver.py 2368 100 16 nn > gpurun_out/prof_nn.log 2>&1; tail -3 gpurun_out/prof_nn.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:selfplay -c 1 -o gpurun_out/prof_r1_ro python scripts/prof_driver.py 2368 100 16 rollout > gpurun_out/prof_ro.log 2>&1; tail -3 gpurun_out/prof_ro.log
ls -la gpurun_out
