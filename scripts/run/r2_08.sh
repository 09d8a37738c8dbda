# round 2, GPU call 8: bench.py sanity on every --config (short), both arms
mkdir -p gpurun_out
for c in 1 0 2 4; do
  extra=""; [ $c = 1 ] && extra="--games-mult 1"
  echo "== ours --config $c"
  timeout 600 python bench.py --config $c --steps 1 --warmup 1 $extra 2> gpurun_out/r2_08_bench_c$c.err | tee gpurun_out/r2_08_bench_c$c.json | cut -c1-1500
  tail -3 gpurun_out/r2_08_bench_c$c.err
done
for c in 1 0; do
  echo "== reference --config $c"
  timeout 600 python bench.py --impl reference --config $c --steps 1 --warmup 1 --cpu-games 256 2> gpurun_out/r2_08_ref_c$c.err | tee gpurun_out/r2_08_ref_c$c.json | cut -c1-600
  tail -3 gpurun_out/r2_08_ref_c$c.err
done
