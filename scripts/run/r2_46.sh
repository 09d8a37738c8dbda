# round 2, GPU call 46: rollout kernel <1024,3>, a Const-only instantiation against the two-way one at HEAD (A/B on one box, bench size)
mkdir -p gpurun_out
for v in "" _rc "" _rc; do
SYN_B200_LIB=$PWD/synthesis_b200/libsynthesis_b200$v.so timeout 600 python scripts/prof_driver.py 909312 800 1 rollout 0 2>&1 | head -1 | sed "s/^/lib$v: /"
done | tee gpurun_out/r2_46_ab.txt
SYN_B200_LIB=$PWD/synthesis_b200/libsynthesis_b200_rc.so timeout 900 python -m pytest tests -x -q -m gpu -k "rollout" 2>&1 | tail -2
