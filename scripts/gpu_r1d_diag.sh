# diagnostics of the thread-per-game kernel: what saturates?  (search mode = one tree per root, short launches)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for T in 8 4 2 1; do timeout 300 python scripts/exp_search.py $T 800 2 2 2>&1 | grep teams; done | tee gpurun_out/diag_teams.log
for E in 100 1600; do timeout 300 python scripts/exp_search.py 8 $E 2 2 2>&1 | grep teams; done | tee -a gpurun_out/diag_teams.log
# traffic + hit rates per thread count
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,smsp__inst_executed.sum,sm__cycles_active.avg
for T in 8 4 2; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:selfplay_nn_tpg -c 1 --csv --log-file gpurun_out/diag_metrics_t$T.csv python scripts/exp_search.py $T 800 2 1 > gpurun_out/diag_metrics_t$T.log 2>&1
done
# one full capture at 8 teams
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:selfplay_nn_tpg -c 1 -o gpurun_out/prof_r1d_tpg8 python scripts/exp_search.py 8 800 2 1 > gpurun_out/prof_r1d.log 2>&1; tail -2 gpurun_out/prof_r1d.log
