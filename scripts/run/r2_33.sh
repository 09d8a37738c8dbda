# round 2, GPU call 33: every warp on its own (selfplay_nn_warp_kernel, SYN_LG_TEAMS=1) against teams of four warps (4)
mkdir -p gpurun_out
SYN_LG_TEAMS=1 timeout 900 python -m pytest tests -x -q -m gpu -k "lanes or mapping or nn or chain or config_sizes" 2>&1 | tail -3 | tee gpurun_out/r2_33_tests.txt
{
export SYN_B200_LIB=$PWD/synthesis_b200/libsynthesis_b200_lgprof.so
for teams in 4 1; do
export SYN_LG_TEAMS=$teams
for gl in 16 32; do
echo "== teams $teams lanes $gl: 16384 NN games, 4096 in flight"; timeout 300 python scripts/prof_driver.py 16384 800 $gl nn 4096 2>&1 | head -3
echo "== teams $teams lanes $gl: 4000 NN games, 1000 in flight"; timeout 300 python scripts/prof_driver.py 4000 800 $gl nn 1000 2>&1 | head -3
done
echo "== teams $teams lanes 16: 10000 NN games in flight"; timeout 300 python scripts/prof_driver.py 10000 800 16 nn 10000 2>&1 | head -1
done
} 2>&1 | tee gpurun_out/r2_33_warp.txt
