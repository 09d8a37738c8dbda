# round-1 session-6 run H: speculative prefetch of the best child's children: parity, A/B throughput
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
{
for rep in 1 2; do
  echo "spec off:"; SYN_TPG_SPEC=0 timeout 300 python scripts/prof_driver.py 454656 800 1 nn 2>&1 | head -1
  echo "spec on:";  SYN_TPG_SPEC=1 timeout 300 python scripts/prof_driver.py 454656 800 1 nn 2>&1 | head -1
done
} | tee gpurun_out/spec_ab.log
