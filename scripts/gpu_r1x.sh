mkdir -p gpurun_out
SYN_TRAIN_PROF=1 timeout 600 python scripts/bench_train.py 1000 200 20000 0 2>&1 | grep -E "train phases|op" | tail -3 | cut -c1-900 | tee gpurun_out/train_prof.log
