# round 2, GPU call 48: rollout kernel, 896 threads per CTA with the Const-only instantiation (72 registers, 20 bytes of spills) against 1,024 (64 registers, 80 bytes)
mkdir -p gpurun_out
L=$PWD/synthesis_b200/libsynthesis_b200_r896.so
for rep in 1 2; do
SYN_B200_LIB=$L timeout 600 python scripts/prof_driver.py 909312 800 1 rollout 0 2>&1 | head -1 | sed "s/^/1024 threads rep $rep: /"
SYN_ROLLOUT_THREADS=896 SYN_B200_LIB=$L timeout 600 python scripts/prof_driver.py 795648 800 1 rollout 0 2>&1 | head -1 | sed "s/^/896 threads rep $rep: /"
done | tee gpurun_out/r2_48_ab.txt
