# full ncu capture of the thread-per-game kernel at HEAD (4 teams, search mode, one wave of 75,776 roots at 800 explores)
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:selfplay_nn_tpg2 -c 1 -o gpurun_out/prof_r1n_tpg2_t4 python scripts/exp_search.py 4 800 1 1 > gpurun_out/prof_r1n.log 2>&1; tail -2 gpurun_out/prof_r1n.log
ls -la gpurun_out/*.ncu-rep
