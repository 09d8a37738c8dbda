# round-1 session-6 run D (2 GPUs): the torchrun bench path at N=2 (weights broadcast + sharded games), then deduplicate re-measured
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err ) 2>&1 | grep real; cat gpurun_out/bench_n2.json | cut -c1-400; tail -3 gpurun_out/bench_n2.err
timeout 300 python -m pytest tests -x -q -m gpu -k "dedup" 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_dedup.log
timeout 300 python scripts/bench_dedup.py 454656 64 5 2>&1 | tail -2 | tee gpurun_out/dedup_454k.log
