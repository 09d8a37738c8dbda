# round-1 session-6 run G: full GPU suite; evaluation sweep, learner and deduplicate measurements for profiles/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 900 python scripts/bench_match.py 1184 800 32 2>&1 | tail -1 | tee gpurun_out/match_bench.log
timeout 600 python scripts/bench_train.py 1000 200 20000 2000 2>&1 | tail -1 | tee gpurun_out/train_bench.log
