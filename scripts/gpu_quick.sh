# quick check of the thread-per-game kernel in search mode (one tree per root): GPU tests, then explores/s per team count
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
{
for rep in 1 2; do for T in ${TEAMS_LIST:-4 5}; do timeout 300 python scripts/exp_search.py $T 800 2 2 2>&1 | grep teams | tail -1; done; done
echo "no reductions:"; SYN_TPG_NO_RED=1 timeout 300 python scripts/exp_search.py 4 800 2 2 2>&1 | grep teams | tail -1
SYN_TPG_PROF=1 timeout 300 python scripts/exp_search.py 4 800 2 1 2>&1 | grep teams
} | tee gpurun_out/quick.log
