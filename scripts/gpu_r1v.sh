mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "train or alpha_zero" 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_train.log
timeout 600 python scripts/bench_train.py 1000 200 20000 2000 2>&1 | tail -1 | tee gpurun_out/train_bench.log
