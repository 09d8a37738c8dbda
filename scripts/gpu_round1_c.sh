mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python scripts/prof_driver.py 75776 100 1 nn
python scripts/prof_driver.py 75776 800 1 nn
python scripts/prof_driver.py 227328 800 1 nn
