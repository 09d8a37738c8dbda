mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
SYN_TPG_TEAMS=8 python scripts/prof_driver.py 151552 800 1 nn
SYN_TPG_TEAMS=8 python scripts/prof_driver.py 454656 800 1 nn
SYN_TPG_TEAMS=4 python scripts/prof_driver.py 227328 800 1 nn
