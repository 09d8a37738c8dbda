# round 2, GPU call 12: auto-calibrated forward chain; split forward with double-buffered TMEM reads
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "nn_eval or chosen_by or close_to_fp32" 2>&1 | grep -v Warning | tail -25 | tee gpurun_out/r2_12_pytest_nn.log
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2_12_pytest.log
{
echo "== auto (random-init weights -> single fp16)"; SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 284160 800 1 nn 2>&1 | head -3
echo "== split forced"; SYN_MLP=split SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 284160 800 1 nn 2>&1 | head -3
echo "== split forced, 4 teams"; SYN_MLP=split SYN_TPG_TEAMS=4 SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 227328 800 1 nn 2>&1 | head -3
} 2>&1 | tee gpurun_out/r2_12_ab.log
