# round 2, GPU call 9 (two GPUs): concatenated shards == one GPU through the library's collectives; bench.py --gpus 2 (short)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -x -q -m gpu -k "two_gpus or communicator" 2>&1 | tail -6 | tee gpurun_out/r2_09_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --games-mult 2 2> gpurun_out/r2_09_bench_n2.err | tee gpurun_out/r2_09_bench_n2.json | cut -c1-400
tail -5 gpurun_out/r2_09_bench_n2.err
timeout 900 python bench.py --gpus 1 --steps 1 --warmup 1 --games-mult 2 --no-cpu-baseline 2> gpurun_out/r2_09_bench_n1.err | tee gpurun_out/r2_09_bench_n1.json | cut -c1-300
