# round 2, GPU call 42: Normal-only kernel instantiations with the cached FPU stream; the tuned kernels are SASS-identical to c36ecb0's
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_gpu_tests_head.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/r2_gpu_tests_head.txt
{
echo "== NN, normal FPU, 189440 games"; timeout 600 python scripts/prof_driver.py 189440 800 1 nn 0 normalfpu 2>&1 | head -1
echo "== rollout, normal FPU, 303104 games"; timeout 600 python scripts/prof_driver.py 303104 800 1 rollout 0 normalfpu 2>&1 | head -1
echo "== NN, const FPU, 568320 games (HEAD)"; timeout 600 python scripts/prof_driver.py 568320 800 1 nn 0 2>&1 | head -1
# (the A/B leg against the build of c36ecb0 needs that library rebuilt from the commit: git archive c36ecb0 | nvcc ... -o synthesis_b200/libsynthesis_b200_c36ecb0.so)
} 2>&1 | tee gpurun_out/r2_42_normal_fpu.txt
timeout 900 python bench.py > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err
python -c "import json; d=json.load(open('gpurun_out/r2_bench_default_n1.json')); print('default', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], round(d['ms_per_step'],1), (d.get('roofline') or {}).get('traffic'))"
