# round 2, GPU call 11: split-fp16 forward (A operand in tensor memory): network tests first, then everything, then A/B vs the single-fp16 chain
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "nn_eval" 2>&1 | tail -25 | tee gpurun_out/r2_11_pytest_nn.log
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/r2_11_pytest.log
{
for rep in 1 2; do
echo "== tpg2 single fp16 (mode 1)"; SYN_MLP=fp16 SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 284160 800 1 nn 2>&1 | head -3
echo "== tpg2 split fp16 (mode 2, default)"; SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 284160 800 1 nn 2>&1 | head -3
done
echo "== split, 4 teams / 6 teams"
SYN_TPG_TEAMS=4 SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 227328 800 1 nn 2>&1 | head -3
SYN_TPG_TEAMS=6 SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 340992 800 1 nn 2>&1 | head -3
} 2>&1 | tee gpurun_out/r2_11_ab.log
