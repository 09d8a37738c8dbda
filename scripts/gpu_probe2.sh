mkdir -p gpurun_out
./scripts/probe/random_sector_probe 32 74 1 | tee gpurun_out/probe_74.log
./scripts/probe/random_sector_probe 32 37 1 | tee gpurun_out/probe_37.log
./scripts/probe/random_sector_probe 32 148 1 | tee gpurun_out/probe_148.log
ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_read_lookup_miss.sum --clock-control none --csv --log-file gpurun_out/probe_ncu.csv ./scripts/probe/random_sector_probe 32 148 1 > /dev/null 2>&1
grep -E "dram__|gpu__time|lts__" gpurun_out/probe_ncu.csv | awk -F'","' '{print $5, $(NF-2), $NF}'
