# round-1 session-6 run F: L2 eviction-priority hints in the thread-per-game kernel: parity, A/B throughput, DRAM traffic
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
{
for rep in 1 2; do
  echo "hints off:"; SYN_TPG_L2HINT=0 timeout 300 python scripts/prof_driver.py 454656 800 1 nn 2>&1 | head -1
  echo "hints on:";  SYN_TPG_L2HINT=1 timeout 300 python scripts/prof_driver.py 454656 800 1 nn 2>&1 | head -1
done
} | tee gpurun_out/l2hint_ab.log
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:selfplay -c 1 --csv --log-file gpurun_out/traffic_l2hint_on.csv python scripts/prof_driver.py 454656 800 1 nn > gpurun_out/traffic_on.log 2>&1; tail -1 gpurun_out/traffic_on.log
grep -E "dram__|lts__|gpu__time" gpurun_out/traffic_l2hint_on.csv | cut -d, -f13-15
