# round 2, GPU call 43: fewer inlined descents in the tuned kernel — A/B on one box (default / no Normal descent / Const-only instantiation)
mkdir -p gpurun_out
for rep in 1 2; do
for v in "" _d1 _d2; do
SYN_B200_LIB=$PWD/synthesis_b200/libsynthesis_b200$v.so timeout 600 python scripts/prof_driver.py 568320 800 1 nn 0 2>&1 | head -1 | sed "s/^/lib$v rep $rep: /"
done; done | tee gpurun_out/r2_43_dispatch_ab.txt
SYN_B200_LIB=$PWD/synthesis_b200/libsynthesis_b200_d2.so timeout 900 python -m pytest tests -x -q -m gpu -k "search or gather or variant or nn" 2>&1 | tail -2
