# round 2, GPU call 51: the whole GPU suite and smoke at the final HEAD
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_gpu_tests_head.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/r2_gpu_tests_head.txt
