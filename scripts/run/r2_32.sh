# round 2, GPU call 32: phase clocks of the split-chain team kernel
mkdir -p gpurun_out
{
export SYN_B200_LIB=$PWD/synthesis_b200/libsynthesis_b200_lgprof.so
export SYN_MLP=split
for gl in 16 32; do
echo "== split chain, lanes $gl: 16384 NN games, 4096 in flight"; timeout 300 python scripts/prof_driver.py 16384 800 $gl nn 4096 2>&1 | head -3
echo "== split chain, lanes $gl: 4000 NN games, 1000 in flight"; timeout 300 python scripts/prof_driver.py 4000 800 $gl nn 1000 2>&1 | head -3
done
} 2>&1 | tee gpurun_out/r2_32_split_clocks.txt
