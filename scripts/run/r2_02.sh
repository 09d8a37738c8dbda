# round 2, GPU call 2: is the kernel DRAM-transaction bound?  Games in flight vs throughput (L2 reuse distance), tpg2 and tpg3
mkdir -p gpurun_out
{
for nf in 94720 75776 56832 47360 37888 28416 18944 9472; do
  g=$((nf*3))
  echo "== tpg3 in flight $nf"; SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py $g 800 1 nn $nf 2>&1 | head -3
done
for nf in 56832 37888 18944; do
  g=$((nf*3))
  echo "== tpg3 CH8 in flight $nf"; SYN_B200_LIB=$PWD/synthesis_b200/libsyn_ch8.so SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py $g 800 1 nn $nf 2>&1 | head -3
done
echo "== tpg3 CH8 in flight 94720"; SYN_B200_LIB=$PWD/synthesis_b200/libsyn_ch8.so SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 284160 800 1 nn 94720 2>&1 | head -3
for t in 4 2 1; do
  nf=$((148*128*t)); g=$((nf*3))
  echo "== tpg2 teams $t in flight $nf"; SYN_TPG_VER=2 SYN_TPG_TEAMS=$t SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py $g 800 1 nn $nf 2>&1 | head -3
done
} 2>&1 | tee gpurun_out/r2_02_inflight.log
# L2 hit rate and DRAM traffic at two operating points (tpg3)
for nf in 94720 37888; do
  g=$((nf*2))
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:selfplay_nn -c 1 --csv --log-file gpurun_out/r2_02_ncu_$nf.csv python scripts/prof_driver.py $g 800 1 nn $nf > gpurun_out/r2_02_ncu_$nf.log 2>&1
  tail -8 gpurun_out/r2_02_ncu_$nf.csv | cut -d, -f5,13,14,15
done
