# round 2, GPU call 14: HEAD test suite (auto mapping) + ncu source profiles of the lane-group kernels at 4,096 games
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2_14_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:selfplay_nn_tc -c 1 -o gpurun_out/r2_14_nn_tc_gl16 python scripts/prof_driver.py 4096 200 16 nn 4096 > gpurun_out/r2_14_ncu_nn.log 2>&1
tail -2 gpurun_out/r2_14_ncu_nn.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:selfplay_rollout_kernel -c 1 -o gpurun_out/r2_14_rollout_gl32 python scripts/prof_driver.py 4096 200 32 rollout 4096 > gpurun_out/r2_14_ncu_ro.log 2>&1
tail -2 gpurun_out/r2_14_ncu_ro.log
ls -la gpurun_out/*.ncu-rep
