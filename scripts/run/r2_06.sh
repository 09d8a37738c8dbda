# round 2, GPU call 6: tpg4 v3 (three-line blocks, wide accesses): full parity suite, A/B against tpg2, ncu counters
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2_06_pytest.log
{
echo "== tpg2 nn 5 teams"; SYN_TPG_VER=2 SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 284160 800 1 nn 2>&1 | head -3
echo "== tpg4 nn 5 teams"; SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 284160 800 1 nn 2>&1 | head -3
echo "== tpg4 nn 5 teams (no prof)"; timeout 300 python scripts/prof_driver.py 284160 800 1 nn 2>&1 | head -2
echo "== tpg4 nn 4 teams"; SYN_TPG_TEAMS=4 SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 227328 800 1 nn 2>&1 | head -3
for nf in 75776 56832 37888 18944; do
  g=$((nf*3))
  echo "== tpg4 in flight $nf"; SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py $g 800 1 nn $nf 2>&1 | head -3
done
echo "== tpg2 rollout 1024"; SYN_TPG_VER=2 timeout 300 python scripts/prof_driver.py 454656 800 1 rollout 2>&1 | head -3
echo "== tpg4 rollout 1024"; timeout 300 python scripts/prof_driver.py 454656 800 1 rollout 2>&1 | head -3
echo "== tpg4 rollout 512"; SYN_ROLLOUT_THREADS=512 timeout 300 python scripts/prof_driver.py 227328 800 1 rollout 2>&1 | head -3
echo "== small batch 4096 games E=800 nn"
timeout 300 python scripts/prof_driver.py 4096 800 1 nn 2>&1 | head -2
} 2>&1 | tee gpurun_out/r2_06_ab.log
for v in 4 2; do
SYN_TPG_VER=$v timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_red.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_local_op_st.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:selfplay_nn -c 1 --csv --log-file gpurun_out/r2_06_ncu_tpg$v.csv python scripts/prof_driver.py 189440 800 1 nn > gpurun_out/r2_06_ncu_tpg$v.log 2>&1
done
