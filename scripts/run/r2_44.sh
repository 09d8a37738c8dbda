# round 2, GPU call 44: rollout and split kernels without the third inlined descent — A/B on one box, then the whole suite
mkdir -p gpurun_out
for rep in 1 2; do
for v in _prev ""; do
SYN_B200_LIB=$PWD/synthesis_b200/libsynthesis_b200$v.so timeout 600 python scripts/prof_driver.py 303104 800 1 rollout 0 2>&1 | head -1 | sed "s/^/lib$v rep $rep: /"
SYN_MLP=split SYN_B200_LIB=$PWD/synthesis_b200/libsynthesis_b200$v.so timeout 600 python scripts/prof_driver.py 189440 800 1 nn 0 2>&1 | head -1 | sed "s/^/split lib$v rep $rep: /"
done; done | tee gpurun_out/r2_44_ab.txt
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_gpu_tests_head.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/r2_gpu_tests_head.txt
