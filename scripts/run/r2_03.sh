# round 2, GPU call 3: full ncu capture (source counters) of selfplay_nn_tpg3_kernel<5,4>, one wave of 94,720 games at 200 explores/move
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:selfplay_nn_tpg3 -c 1 -o gpurun_out/r2_03_tpg3_t5 python scripts/prof_driver.py 94720 200 1 nn > gpurun_out/r2_03_ncu.log 2>&1
tail -3 gpurun_out/r2_03_ncu.log
ls -la gpurun_out/*.ncu-rep
