# round 2, GPU call 21: lane-group network kernel with the in-place A tile, two 256-thread CTAs per SM and even seating
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -k "not two_gpus" 2>&1 | tail -8 | tee gpurun_out/r2_21_tests.txt
{
for gl in 16 32; do
echo "== lanes $gl: 4096 NN games E=800"; timeout 300 python scripts/prof_driver.py 4096 800 $gl nn 4096 2>&1 | head -3
echo "== lanes $gl: 1000 NN games E=800"; timeout 300 python scripts/prof_driver.py 1000 800 $gl nn 1000 2>&1 | head -3
done
echo "== lanes 32: 256 rollout games"; timeout 300 python scripts/prof_driver.py 256 800 32 rollout 256 2>&1 | head -2
for c in 0 2; do
echo "== bench --config $c"; timeout 600 python bench.py --config $c 2>gpurun_out/r2_21_bench_c$c.err > gpurun_out/r2_21_bench_c$c.json; python -c "import json,sys; d=json.load(open('gpurun_out/r2_21_bench_c$c.json')); print(d['value'], d['e2e'], d['cpu_baseline']['value'], d['ms_per_step'], d.get('wall_ms_per_step'))"
done
} 2>&1 | tee gpurun_out/r2_21_nn.txt
