# round 2, GPU call 35: one training iteration of the reference as shipped, every stage on the GPU, phases timed (scripts/bench_iteration.py)
mkdir -p gpurun_out
timeout 900 python scripts/bench_iteration.py 3 1000 2> gpurun_out/r2_iteration.err | tee gpurun_out/r2_iteration.json | cut -c1-1500
tail -3 gpurun_out/r2_iteration.err
