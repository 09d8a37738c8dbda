# round-1 session-6 run I: L2 persisting carve-out for the evict_last lines: A/B throughput + DRAM traffic
mkdir -p gpurun_out
{
for rep in 1 2; do
  echo "persist 0 MB:"; SYN_L2_PERSIST_MB=0 timeout 300 python scripts/prof_driver.py 454656 800 1 nn 2>&1 | head -1
  echo "persist max:";  SYN_L2_PERSIST_MB=4096 timeout 300 python scripts/prof_driver.py 454656 800 1 nn 2>&1 | head -1
done
echo "persist max, hints off:"; SYN_L2_PERSIST_MB=4096 SYN_TPG_L2HINT=0 timeout 300 python scripts/prof_driver.py 454656 800 1 nn 2>&1 | head -1
} | tee gpurun_out/persist_ab.log
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:selfplay -c 1 --csv --log-file gpurun_out/traffic_persist.csv python scripts/prof_driver.py 454656 800 1 nn > gpurun_out/traffic_persist.log 2>&1
grep -E "dram__|lts__|gpu__time" gpurun_out/traffic_persist.csv | awk -F'","' '{print $13, $15}'
