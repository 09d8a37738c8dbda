# round 2, GPU call 22: lane-group network kernel, one 512-thread CTA per SM against two 256-thread CTAs, with the seats
# refilled (4 x the games in flight) so that the tail of a single wave does not decide the number
mkdir -p gpurun_out
{
for v in lgprof lgprof512; do
export SYN_B200_LIB=$PWD/synthesis_b200/libsynthesis_b200_$v.so
for gl in 16 32; do
echo "== $v lanes $gl: 16384 NN games, 4096 in flight"; timeout 300 python scripts/prof_driver.py 16384 800 $gl nn 4096 2>&1 | head -3
echo "== $v lanes $gl: 4096 NN games, 4096 in flight"; timeout 300 python scripts/prof_driver.py 4096 800 $gl nn 4096 2>&1 | head -3
echo "== $v lanes $gl: 4000 NN games, 1000 in flight"; timeout 300 python scripts/prof_driver.py 4000 800 $gl nn 1000 2>&1 | head -3
done
done
} 2>&1 | tee gpurun_out/r2_22_nn_ab.txt
