# round 2, GPU call 36: the FPU stream's key and block kept in shared memory (lane-group kernels); the shipped iteration again
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "normal_fpu or fpu or lanes or mapping or variant" 2>&1 | tail -3 | tee gpurun_out/r2_36_tests.txt
timeout 900 python scripts/bench_iteration.py 3 1000 2> gpurun_out/r2_iteration.err > gpurun_out/r2_iteration.json
python -c "
import json; d=json.load(open('gpurun_out/r2_iteration.json'))
for it in d['iterations']: print({k: (round(v,3) if isinstance(v,float) else v) for k,v in it.items() if k!='last_epoch_loss'})
print(d['host_estimate_for_last_iteration']); print('speedup', d['speedup_last_iteration'])"
