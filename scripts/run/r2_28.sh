# round 2, GPU call 28: where the lane-group kernels stop winning (crossover against a thread per game), split chain at small batches
mkdir -p gpurun_out
{
for n in 9000 20000 40000; do
for gl in 32 1; do
echo "== rollout lanes $gl: $n games"; timeout 300 python scripts/prof_driver.py $n 800 $gl rollout $n 2>&1 | head -1
done; done
for n in 6000 10000 16000; do
for gl in 16 1; do
echo "== NN lanes $gl: $n games in flight $n"; timeout 300 python scripts/prof_driver.py $n 800 $gl nn $n 2>&1 | head -1
done; done
export SYN_MLP=split
for gl in 16 32 1; do
echo "== split chain, NN lanes $gl: 16384 games, 4096 in flight"; timeout 300 python scripts/prof_driver.py 16384 800 $gl nn 4096 2>&1 | head -1
echo "== split chain, NN lanes $gl: 4000 games, 1000 in flight"; timeout 300 python scripts/prof_driver.py 4000 800 $gl nn 1000 2>&1 | head -1
done
} 2>&1 | tee gpurun_out/r2_28_crossover.txt
