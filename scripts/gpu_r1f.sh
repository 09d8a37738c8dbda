mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
{
for T in 8 6 4 2; do SYN_TPG_TEAMS=$T timeout 300 python scripts/exp_search.py $T 800 2 2 2>&1 | grep teams; done
SYN_TPG_PROF=1 timeout 300 python scripts/exp_search.py 8 800 2 1 2>&1 | grep teams
SYN_TPG_PROF=1 timeout 300 python scripts/exp_search.py 6 800 2 1 2>&1 | grep teams
SYN_TPG_PROF=1 timeout 300 python scripts/exp_search.py 4 800 2 1 2>&1 | grep teams
} | tee gpurun_out/tpg2b_teams.log
( time timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_t8.json 2> gpurun_out/bench_t8.err ) 2>&1 | grep real; cat gpurun_out/bench_t8.json; tail -3 gpurun_out/bench_t8.err
( time SYN_TPG_TEAMS=4 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_t4.json 2> gpurun_out/bench_t4.err ) 2>&1 | grep real; cat gpurun_out/bench_t4.json; tail -3 gpurun_out/bench_t4.err
