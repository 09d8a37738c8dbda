mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "nn or group" 2>&1 | tail -30
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
for gl in 16 32; do
  timeout 600 python bench.py --steps 2 --warmup 3 --group-lanes $gl --no-cpu-baseline > gpurun_out/bench_tc_gl$gl.json 2> gpurun_out/bench_tc_gl$gl.err; python -c "
import json;d=json.load(open('gpurun_out/bench_tc_gl$gl.json'));print('GL$gl value',d['value']/1e6,'e2e',d['e2e']['value']/1e6,'ms',d['ms_per_step'])"; tail -3 gpurun_out/bench_tc_gl$gl.err
done
