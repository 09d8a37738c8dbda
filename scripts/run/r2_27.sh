# round 2, GPU call 27 (--gpus 2): the two-GPU equality test and the N=2 bench after broadcast_weights learned to calibrate the forward chain
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "two_gpus or communicator or lanes or mapping or occupy" 2>&1 | tail -4 | tee gpurun_out/r2_27_tests.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_default_n2.json 2> gpurun_out/r2_bench_default_n2.err
python -c "import json; d=json.load(open('gpurun_out/r2_bench_default_n2.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['dtype'][:90])"
