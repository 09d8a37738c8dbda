mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for rep in 1 2; do SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 454656 800 1 nn 2>&1 | head -2; done | tee gpurun_out/fresh_leaf.log
