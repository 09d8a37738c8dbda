mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "train or dedup or alpha_zero" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_sub.log
timeout 600 python scripts/bench_dedup.py 454656 64 5 2>&1 | tail -1 | tee gpurun_out/dedup_454k.log
SYN_TRAIN_PROF=1 timeout 600 python scripts/bench_train.py 1000 200 20000 2000 2>&1 | grep -E "train phases|op" | tail -2 | tee gpurun_out/train_bench.log
