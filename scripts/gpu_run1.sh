nvidia-smi --query-gpu=name,memory.total --format=csv
nproc
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -40
