mkdir -p gpurun_out
python scripts/prof_driver.py 2048 200 32
timeout 900 ncu --set full --clock-control none --import-source on -k regex:selfplay -c 2 -o gpurun_out/prof_r1_gl32 python scripts/prof_driver.py 2048 200 32 > gpurun_out/prof1.log 2>&1; tail -5 gpurun_out/prof1.log
ls -la gpurun_out
