# round 2, GPU call 38: Normal FPU with the stream's key and block cached in the slot record (thread-per-game kernels)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -k "fpu or variant or layouts or config or search" 2>&1 | tail -3 | tee gpurun_out/r2_38_tests.txt
{
echo "== NN, normal FPU, 189440 games"; timeout 600 python scripts/prof_driver.py 189440 800 1 nn 0 normalfpu 2>&1 | head -1
echo "== rollout, normal FPU, 303104 games"; timeout 600 python scripts/prof_driver.py 303104 800 1 rollout 0 normalfpu 2>&1 | head -1
echo "== NN, const FPU, 189440 games"; timeout 600 python scripts/prof_driver.py 189440 800 1 nn 0 2>&1 | head -1
} 2>&1 | tee gpurun_out/r2_38_normal_fpu.txt
