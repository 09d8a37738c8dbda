# round 2, GPU call 37: what the shipped Normal(1.0, 0.1) first-play urgency costs at full scale (thread per game)
mkdir -p gpurun_out
{
echo "== NN, const FPU, 189440 games"; timeout 600 python scripts/prof_driver.py 189440 800 1 nn 0 2>&1 | head -2
echo "== NN, normal FPU, 189440 games"; timeout 600 python scripts/prof_driver.py 189440 800 1 nn 0 normalfpu 2>&1 | head -2
echo "== rollout, const FPU, 303104 games"; timeout 600 python scripts/prof_driver.py 303104 800 1 rollout 0 2>&1 | head -1
echo "== rollout, normal FPU, 303104 games"; timeout 600 python scripts/prof_driver.py 303104 800 1 rollout 0 normalfpu 2>&1 | head -1
} 2>&1 | tee gpurun_out/r2_37_normal_fpu.txt
