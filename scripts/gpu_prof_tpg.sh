mkdir -p gpurun_out
# one full capture of the thread-per-game kernel on a short workload (E=100), all SMs busy
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:selfplay -c 1 -o gpurun_out/prof_r1_tpg python scripts/prof_driver.py 75776 100 1 nn > gpurun_out/prof_tpg.log 2>&1; tail -3 gpurun_out/prof_tpg.log
python scripts/prof_driver.py 75776 100 1 nn
python scripts/prof_driver.py 75776 800 1 nn
