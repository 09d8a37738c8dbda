# round 2, GPU call 20: windowed playouts in the lane-group rollout kernel — parity first, then the small-batch numbers
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -k "rollout or lanes or mapping or config_sizes or smoke" 2>&1 | tail -8 | tee gpurun_out/r2_20_tests.txt
{
for gl in 16 32; do
echo "== lanes $gl: 256 rollout games E=800"; timeout 300 python scripts/prof_driver.py 256 800 $gl rollout 256 2>&1 | head -3
echo "== lanes $gl: 4096 rollout games E=800"; timeout 300 python scripts/prof_driver.py 4096 800 $gl rollout 4096 2>&1 | head -3
done
echo "== bench --config 0"; timeout 600 python bench.py --config 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['cpu_baseline']['value'])"
} 2>&1 | tee gpurun_out/r2_20_rollout.txt
