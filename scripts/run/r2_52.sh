# round 2, GPU call 52 (--gpus 2): N=2 bench at the final HEAD
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_default_n2.json 2> gpurun_out/r2_bench_default_n2.err
python -c "import json; d=json.load(open('gpurun_out/r2_bench_default_n2.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['dtype'][:70])"
