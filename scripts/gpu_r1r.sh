# round-1 session-6 run E: full GPU suite (incl. the learner), learner throughput vs PyTorch CPU
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_train.py 1000 200 20000 2000 2>&1 | tail -2 | tee gpurun_out/train_bench.log
