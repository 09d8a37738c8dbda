# round 2, GPU call 1: full parity suite on the tpg3 build, then tpg2 / tpg3 A-B at the bench configuration
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2_01_pytest.log
ab() { echo "== $*"; env "$@" ; }
{
for rep in 1 2; do
  echo "== tpg2 nn 5 teams"; SYN_TPG_VER=2 timeout 300 python scripts/prof_driver.py 284160 800 1 nn 2>&1 | head -3
  echo "== tpg3 nn 5 teams"; timeout 300 python scripts/prof_driver.py 284160 800 1 nn 2>&1 | head -3
done
echo "== tpg3 nn 4 teams"; SYN_TPG_TEAMS=4 timeout 300 python scripts/prof_driver.py 227328 800 1 nn 2>&1 | head -3
echo "== tpg3 nn 6 teams"; SYN_TPG_TEAMS=6 timeout 300 python scripts/prof_driver.py 340992 800 1 nn 2>&1 | head -3
echo "== tpg3 nn 5 teams PROF"; SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 284160 800 1 nn 2>&1 | head -3
echo "== tpg2 nn 5 teams PROF"; SYN_TPG_VER=2 SYN_TPG_PROF=1 timeout 300 python scripts/prof_driver.py 284160 800 1 nn 2>&1 | head -3
echo "== tpg2 rollout 1024"; SYN_TPG_VER=2 timeout 300 python scripts/prof_driver.py 454656 800 1 rollout 2>&1 | head -3
echo "== tpg3 rollout 1024"; timeout 300 python scripts/prof_driver.py 454656 800 1 rollout 2>&1 | head -3
echo "== tpg3 rollout 768"; SYN_ROLLOUT_THREADS=768 timeout 300 python scripts/prof_driver.py 340992 800 1 rollout 2>&1 | head -3
echo "== tpg3 rollout 512"; SYN_ROLLOUT_THREADS=512 timeout 300 python scripts/prof_driver.py 227328 800 1 rollout 2>&1 | head -3
echo "== small batch: 4096 games E=800, tpg2 then tpg3"
SYN_TPG_VER=2 timeout 300 python scripts/prof_driver.py 4096 800 1 nn 2>&1 | head -3
timeout 300 python scripts/prof_driver.py 4096 800 1 nn 2>&1 | head -3
timeout 300 python scripts/prof_driver.py 4096 800 1 rollout 2>&1 | head -3
timeout 300 python scripts/prof_driver.py 1000 800 1 nn 2>&1 | head -3
} 2>&1 | tee gpurun_out/r2_01_ab.log
