# round 2, GPU call 34: the record at HEAD — whole GPU suite, smoke, bench.py on every config (both arms), launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_gpu_tests_head.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/r2_gpu_tests_head.txt
timeout 900 python bench.py > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err
for c in 0 2 4; do
  timeout 900 python bench.py --config $c > gpurun_out/r2_bench_config${c}_n1.json 2> gpurun_out/r2_bench_config${c}_n1.err
done
timeout 900 python bench.py --leaf rollout --steps 1 --warmup 1 > gpurun_out/r2_bench_rollout_n1.json 2> gpurun_out/r2_bench_rollout_n1.err
for c in 1 0 2 4; do
  timeout 900 python bench.py --impl reference --config $c --steps 1 --warmup 1 > gpurun_out/r2_bench_reference_config${c}.json 2> gpurun_out/r2_bench_reference_config${c}.err
done
for f in default_n1 config0_n1 config2_n1 config4_n1 rollout_n1 reference_config1 reference_config0 reference_config2 reference_config4; do
python -c "import json; d=json.load(open('gpurun_out/r2_bench_$f.json')); print('$f', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], round(d['ms_per_step'],1), (d.get('roofline') or {}).get('traffic'))"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_n1.csv python bench.py --steps 1 --warmup 1 --games-mult 1 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_config2.csv python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_launches_bench_c2.log 2>&1
grep -c selfplay gpurun_out/r2_launches_bench_n1.csv gpurun_out/r2_launches_bench_config2.csv
