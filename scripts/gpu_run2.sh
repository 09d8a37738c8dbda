mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|^CPU\(s\)"
for gl in 16 32; do
  timeout 600 python bench.py --steps 3 --warmup 3 --group-lanes $gl --no-cpu-baseline > gpurun_out/bench_nn_gl$gl.json 2> gpurun_out/bench_nn_gl$gl.err; tail -c 1500 gpurun_out/bench_nn_gl$gl.json; tail -3 gpurun_out/bench_nn_gl$gl.err
  timeout 600 python bench.py --steps 3 --warmup 3 --group-lanes $gl --leaf rollout --no-cpu-baseline > gpurun_out/bench_ro_gl$gl.json 2> gpurun_out/bench_ro_gl$gl.err; tail -c 1500 gpurun_out/bench_ro_gl$gl.json; tail -3 gpurun_out/bench_ro_gl$gl.err
done
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>&1; cat gpurun_out/bench_reference.json
