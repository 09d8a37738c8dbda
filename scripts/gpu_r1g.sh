mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
{
SYN_TPG_V=2 timeout 300 python scripts/exp_search.py 8 800 2 1 2>&1 | grep teams
SYN_TPG_CW=3 timeout 300 python scripts/exp_search.py 8 800 2 2 2>&1 | grep teams
SYN_TPG_CW=5 timeout 300 python scripts/exp_search.py 8 800 2 2 2>&1 | grep teams
timeout 300 python scripts/exp_search.py 6 800 2 2 2>&1 | grep teams
SYN_TPG_CW=5 timeout 300 python scripts/exp_search.py 4 800 2 2 2>&1 | grep teams
SYN_TPG_CW=9 timeout 300 python scripts/exp_search.py 4 800 2 2 2>&1 | grep teams
timeout 300 python scripts/exp_search.py 2 800 2 2 2>&1 | grep teams
SYN_TPG_PROF=1 timeout 300 python scripts/exp_search.py 8 800 2 1 2>&1 | grep teams
SYN_TPG_PROF=1 timeout 300 python scripts/exp_search.py 4 800 2 1 2>&1 | grep teams
SYN_TPG_CW=5 timeout 300 python scripts/exp_search.py 8 1600 2 1 2>&1 | grep teams
} | tee gpurun_out/tpg3_teams.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_sectors.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.per_cycle_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
for T in 8 4; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:selfplay_nn_tpg -c 1 --csv --log-file gpurun_out/tpg3_metrics_t$T.csv python scripts/exp_search.py $T 800 2 1 > gpurun_out/tpg3_metrics_t$T.log 2>&1
done
