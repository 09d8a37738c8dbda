# round 2, GPU call 16: HEAD test suite (automatic mapping) + ncu source summaries of the lane-group kernels at 4,096 games
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -X faulthandler -m pytest tests -x -q -m gpu 2>&1 | grep -v "^  File \"/opt" | tail -12 | tee gpurun_out/r2_16_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:selfplay_nn_tc -c 1 -o /tmp/ncu/nn_tc_gl16 python scripts/prof_driver.py 4096 200 16 nn 4096 > gpurun_out/r2_16_ncu_nn.log 2>&1
python scripts/ncu_source_summary.py /tmp/ncu/nn_tc_gl16.ncu-rep 45 > gpurun_out/r2_16_nn_tc_gl16_source_summary.txt 2>&1
ncu -i /tmp/ncu/nn_tc_gl16.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; vals=rows[2] if len(rows)>2 else rows[1]
keep=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__grid_size','launch__block_size']
for k,v in zip(hdr,vals):
    if k in keep: print(k,v)
" > gpurun_out/r2_16_nn_tc_gl16_raw.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:selfplay_rollout_kernel -c 1 -o /tmp/ncu/rollout_gl32 python scripts/prof_driver.py 4096 200 32 rollout 4096 > gpurun_out/r2_16_ncu_ro.log 2>&1
python scripts/ncu_source_summary.py /tmp/ncu/rollout_gl32.ncu-rep 45 > gpurun_out/r2_16_rollout_gl32_source_summary.txt 2>&1
tail -2 gpurun_out/r2_16_ncu_nn.log gpurun_out/r2_16_ncu_ro.log
