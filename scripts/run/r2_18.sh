# round 2, GPU call 18 (--gpus 8): the scaling run the driver does at round end, both arms, through the in-library NCCL path
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r2_bench_default_n$N.json 2> gpurun_out/r2_bench_default_n$N.err
tail -c 1500 gpurun_out/r2_bench_default_n$N.json; tail -5 gpurun_out/r2_bench_default_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/r2_bench_reference_n$N.json 2> gpurun_out/r2_bench_reference_n$N.err
tail -c 600 gpurun_out/r2_bench_reference_n$N.json
