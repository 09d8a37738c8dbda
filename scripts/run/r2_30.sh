# round 2, GPU call 30: learner with batch sizes 64 / 160 against PyTorch; per-step trace of the e2e leg of bench.py --config 0
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "train" -s 2>&1 | grep -v Warning | tail -14 | tee gpurun_out/r2_30_train_tests.txt
SYN_BENCH_TRACE=1 timeout 600 python bench.py --config 0 --steps 12 --no-cpu-baseline 2> gpurun_out/r2_30_c0_trace.txt | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'])"
grep "e2e step" gpurun_out/r2_30_c0_trace.txt
timeout 300 python scripts/bench_train.py 2>&1 | tail -6 | tee gpurun_out/r2_30_train_bench.txt
