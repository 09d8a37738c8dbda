# round 2, GPU call 26: whole GPU suite at HEAD (windowed playouts, lane groups in teams) and the small-config bench lines
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2_26_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/r2_26_tests.txt
for c in 0 2; do
timeout 900 python bench.py --config $c > gpurun_out/r2_bench_config${c}_n1.json 2> gpurun_out/r2_bench_config${c}_n1.err
python -c "import json; d=json.load(open('gpurun_out/r2_bench_config${c}_n1.json')); print($c, d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['ms_per_step'], d['roofline']['kernel'])"
done
timeout 900 python bench.py --impl reference --config 2 --steps 1 --warmup 1 > gpurun_out/r2_bench_reference_config2.json 2>/dev/null
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:selfplay -c 1 --csv --log-file gpurun_out/r2_traffic_nn_e1600_g16384.csv python scripts/prof_driver.py 16384 1600 0 nn 4096 > gpurun_out/r2_traffic_c2.log 2>&1
tail -2 gpurun_out/r2_traffic_c2.log
