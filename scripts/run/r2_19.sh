# round 2, GPU call 19: where a round goes in the lane-group kernels at small batches (the -DSYN_LG_PROF build), and
# how many DRAM bytes one random sector read moves (ncu on the probe)
mkdir -p gpurun_out
{
export SYN_B200_LIB=$PWD/synthesis_b200/libsynthesis_b200_lgprof.so
for gl in 16 32; do
echo "== lanes $gl: 4096 NN games E=800"; timeout 300 python scripts/prof_driver.py 4096 800 $gl nn 4096 2>&1 | head -3
echo "== lanes $gl: 1000 NN games E=800"; timeout 300 python scripts/prof_driver.py 1000 800 $gl nn 1000 2>&1 | head -3
echo "== lanes $gl: 256 rollout games E=800"; timeout 300 python scripts/prof_driver.py 256 800 $gl rollout 256 2>&1 | head -3
echo "== lanes $gl: 4096 rollout games E=800"; timeout 300 python scripts/prof_driver.py 4096 800 $gl rollout 4096 2>&1 | head -3
done
unset SYN_B200_LIB
} 2>&1 | tee gpurun_out/r2_19_lgprof.txt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/family_read_probe scripts/probe/family_read_probe.cu
timeout 600 ncu --metrics dram__bytes_read.sum,lts__t_sectors_op_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:probe -c 10 --csv --log-file gpurun_out/r2_19_probe_ncu.csv /tmp/family_read_probe 8 > gpurun_out/r2_19_probe.log 2>&1
tail -3 gpurun_out/r2_19_probe.log
