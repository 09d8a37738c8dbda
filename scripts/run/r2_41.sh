# round 2, GPU call 41: A/B on ONE box — the library at HEAD against the build of c36ecb0 (1.500 G in its record run), default bench configuration
mkdir -p gpurun_out
for rep in 1 2; do
for v in "" _c36ecb0; do
SYN_B200_LIB=$PWD/synthesis_b200/libsynthesis_b200$v.so timeout 600 python scripts/prof_driver.py 568320 800 1 nn 0 2>&1 | head -1 | sed "s/^/lib$v rep $rep: /"
done; done | tee gpurun_out/r2_41_ab.txt
