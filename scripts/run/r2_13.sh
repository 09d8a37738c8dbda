# round 2, GPU call 13: the lane-group kernels (tree.cuh / selfplay.cuh) at SMALL batches, beside the seated thread-per-game kernels
mkdir -p gpurun_out
{
for gl in 1 16 32; do
echo "== group lanes $gl: 4096 games E=800 nn (in flight 4096)"; timeout 300 python scripts/prof_driver.py 4096 800 $gl nn 4096 2>&1 | head -2
echo "== group lanes $gl: 4096 games E=800 rollout"; timeout 300 python scripts/prof_driver.py 4096 800 $gl rollout 4096 2>&1 | head -2
echo "== group lanes $gl: 256 games E=800 rollout"; timeout 300 python scripts/prof_driver.py 256 800 $gl rollout 256 2>&1 | head -2
echo "== group lanes $gl: 1000 games E=800 nn"; timeout 300 python scripts/prof_driver.py 1000 800 $gl nn 1000 2>&1 | head -2
done
} 2>&1 | grep -v "phase share" | tee gpurun_out/r2_13_small.log
