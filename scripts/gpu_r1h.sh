mkdir -p gpurun_out
{
SYN_TPG_V=2 SYN_L2_FETCH=32 timeout 300 python scripts/exp_search.py 8 800 2 2 2>&1 | grep teams
SYN_TPG_V=2 SYN_L2_FETCH=128 timeout 300 python scripts/exp_search.py 8 800 2 2 2>&1 | grep teams
SYN_TPG_CW=3 SYN_L2_FETCH=32 timeout 300 python scripts/exp_search.py 8 800 2 2 2>&1 | grep teams
} | tee gpurun_out/l2fetch.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_write_lookup_miss.sum,l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum,smsp__sass_inst_executed_op_global_ld.sum,smsp__sass_inst_executed_op_global_st.sum
SYN_TPG_V=2 SYN_L2_FETCH=32 timeout 600 ncu --metrics $M --clock-control none -k regex:selfplay_nn_tpg -c 1 --csv --log-file gpurun_out/m_v2_f32.csv python scripts/exp_search.py 8 800 2 1 > gpurun_out/m_v2_f32.log 2>&1
SYN_TPG_V=2 timeout 600 ncu --metrics $M --clock-control none -k regex:selfplay_nn_tpg -c 1 --csv --log-file gpurun_out/m_v2.csv python scripts/exp_search.py 8 800 2 1 > gpurun_out/m_v2.log 2>&1
SYN_TPG_CW=3 timeout 600 ncu --metrics $M --clock-control none -k regex:selfplay_nn_tpg -c 1 --csv --log-file gpurun_out/m_v3.csv python scripts/exp_search.py 8 800 2 1 > gpurun_out/m_v3.log 2>&1
# full captures (source-level) of the split-record kernel at 8 and 4 teams
SYN_TPG_CW=3 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:selfplay_nn_tpg -c 1 -o gpurun_out/prof_r1h_tpg3_t8 python scripts/exp_search.py 8 800 2 1 > gpurun_out/prof_r1h_t8.log 2>&1
SYN_TPG_CW=5 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:selfplay_nn_tpg -c 1 -o gpurun_out/prof_r1h_tpg3_t4 python scripts/exp_search.py 4 800 2 1 > gpurun_out/prof_r1h_t4.log 2>&1
ls -la gpurun_out
