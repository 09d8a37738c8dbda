# round 2, GPU call 24: lane groups in teams that share ONE A tile (107 KB of shared memory, the rest stays L1)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -k "lanes or mapping or config_sizes or nn or chain or smoke" 2>&1 | tail -6 | tee gpurun_out/r2_24_tests.txt
{
export SYN_B200_LIB=$PWD/synthesis_b200/libsynthesis_b200_lgprof.so
for teams in 4 5 6; do
export SYN_LG_TEAMS=$teams
for gl in 16 32; do
echo "== teams $teams lanes $gl: 16384 NN games, 4096 in flight"; timeout 300 python scripts/prof_driver.py 16384 800 $gl nn 4096 2>&1 | head -3
echo "== teams $teams lanes $gl: 4000 NN games, 1000 in flight"; timeout 300 python scripts/prof_driver.py 4000 800 $gl nn 1000 2>&1 | head -3
done
done
} 2>&1 | tee gpurun_out/r2_24_teams.txt
