# round 2, GPU call 47: the record at HEAD after the rollout Const-only instantiation — suite, smoke, default / rollout / small-config benches
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_gpu_tests_head.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/r2_gpu_tests_head.txt
timeout 900 python bench.py > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err
timeout 900 python bench.py --leaf rollout --steps 1 --warmup 1 > gpurun_out/r2_bench_rollout_n1.json 2> gpurun_out/r2_bench_rollout_n1.err
for c in 0 2; do timeout 900 python bench.py --config $c > gpurun_out/r2_bench_config${c}_n1.json 2> gpurun_out/r2_bench_config${c}_n1.err; done
for f in default_n1 rollout_n1 config0_n1 config2_n1; do
python -c "import json; d=json.load(open('gpurun_out/r2_bench_$f.json')); print('$f', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], round(d['ms_per_step'],1), (d.get('roofline') or {}).get('traffic'), '%.3f' % d['roofline']['frac'])"
done
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:selfplay -c 1 --csv --log-file gpurun_out/r2_traffic_rollout_e800_g909312.csv python scripts/prof_driver.py 909312 800 0 rollout > gpurun_out/r2_traffic_rollout.log 2>&1
tail -1 gpurun_out/r2_traffic_rollout.log | cut -c1-200
