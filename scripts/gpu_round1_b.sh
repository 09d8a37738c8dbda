mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for teams in 4 2; do
SYN_TPG_TEAMS=$teams timeout 900 python bench.py --steps 2 --warmup 1 --group-lanes 1 --games $((148*128*teams)) --no-cpu-baseline > gpurun_out/bench_tpg_t$teams.json 2> gpurun_out/bench_tpg_t$teams.err; python -c "
import json;d=json.load(open('gpurun_out/bench_tpg_t$teams.json'));print('teams $teams value',d['value']/1e6,'e2e',d['e2e']['value']/1e6,'ms',d['ms_per_step'], d['clocks'])"; tail -3 gpurun_out/bench_tpg_t$teams.err
done
