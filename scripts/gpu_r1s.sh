mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "train" -s 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_train.log
