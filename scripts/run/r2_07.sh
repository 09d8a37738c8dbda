# round 2, GPU call 7: product kernels (tpg2 + seating) full suite incl. the communicator-of-one tests; A/B of constant-bank MLP biases
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2_07_pytest.log
{
for rep in 1 2; do
echo "== tpg2 nn 5 teams"; SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 284160 800 1 nn 2>&1 | head -3
echo "== tpg2 nn 5 teams, constant-bank biases"; SYN_B200_LIB=$PWD/synthesis_b200/libsyn_cbias.so SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 284160 800 1 nn 2>&1 | head -3
done
echo "== small batches (seated over all SMs): 4096 games E=1600 nn, 4096 E=800 nn, 1000 E=800 nn, 256 E=800 rollout, 4096 E=800 rollout"
timeout 300 python scripts/prof_driver.py 4096 1600 1 nn 2>&1 | head -2
SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 4096 800 1 nn 2>&1 | head -3
SYN_TPG_TEAMS=1 SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 4096 800 1 nn 4096 2>&1 | head -3
timeout 300 python scripts/prof_driver.py 1000 800 1 nn 2>&1 | head -2
timeout 300 python scripts/prof_driver.py 256 800 1 rollout 2>&1 | head -2
timeout 300 python scripts/prof_driver.py 4096 800 1 rollout 2>&1 | head -2
} 2>&1 | tee gpurun_out/r2_07_ab.log
