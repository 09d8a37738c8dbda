mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "dedup" 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_sub.log
timeout 600 python scripts/bench_dedup.py 454656 64 5 2>&1 | tail -1 | tee gpurun_out/dedup_454k.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:insert_kernel|rep_kernel|rs_|scan_|heads_|starts_|reduce_|fill_kernel|gather_kernel" --csv --log-file gpurun_out/launches_dedup.csv python scripts/bench_dedup.py 454656 64 1 > gpurun_out/launches_dedup.log 2>&1
