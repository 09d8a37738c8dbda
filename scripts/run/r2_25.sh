# round 2, GPU call 25: team kernel with up to four TMEM reads in flight per wait; speculative L2 prefetch of every child's family A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -k "lanes or mapping or nn or chain" 2>&1 | tail -3 | tee gpurun_out/r2_25_tests.txt
{
for v in lgprof lgprof_pf; do
export SYN_B200_LIB=$PWD/synthesis_b200/libsynthesis_b200_$v.so
for gl in 16 32; do
echo "== $v teams 4 lanes $gl: 16384 NN games, 4096 in flight"; timeout 300 python scripts/prof_driver.py 16384 800 $gl nn 4096 2>&1 | head -3
echo "== $v teams 4 lanes $gl: 4000 NN games, 1000 in flight"; timeout 300 python scripts/prof_driver.py 4000 800 $gl nn 1000 2>&1 | head -3
done
echo "== $v lanes 32: 256 rollout games"; timeout 300 python scripts/prof_driver.py 256 800 32 rollout 256 2>&1 | head -3
echo "== $v lanes 32: 4096 rollout games"; timeout 300 python scripts/prof_driver.py 4096 800 32 rollout 4096 2>&1 | head -3
done
} 2>&1 | tee gpurun_out/r2_25_teams.txt
