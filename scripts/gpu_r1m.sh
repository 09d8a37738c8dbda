# round-1 session-5 run A: GPU tests, sector reach probe, one full ncu capture of the thread-per-game kernel (4 teams, search mode)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 120 ./scripts/probe/sector_reach_probe 32 | tee gpurun_out/sector_reach_probe.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:selfplay_nn_tpg2 -c 1 -o gpurun_out/prof_r1m_tpg2_t4 python scripts/exp_search.py 4 800 1 1 > gpurun_out/prof_r1m.log 2>&1; tail -2 gpurun_out/prof_r1m.log
ls -la gpurun_out/*.ncu-rep
