# round 2, GPU call 39: whole GPU suite, smoke and the default bench at HEAD after the FPU-stream change
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_gpu_tests_head.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/r2_gpu_tests_head.txt
timeout 900 python bench.py > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err
for c in 0 2; do timeout 900 python bench.py --config $c --no-cpu-baseline > gpurun_out/r2_39_c$c.json 2>/dev/null; done
for f in r2_bench_default_n1 r2_39_c0 r2_39_c2; do
python -c "import json; d=json.load(open('gpurun_out/$f.json')); print('$f', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], round(d['ms_per_step'],1), (d.get('roofline') or {}).get('traffic'))"
done
