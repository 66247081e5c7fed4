# Large traces on one GPU: complete proofs at 2^23 and 2^24 rows (device-resident inputs)
set -x
mkdir -p gpurun_out
for L in 23 24; do
  timeout 900 python bench.py --log-rows $L --quick --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/quick_log$L.json 2> gpurun_out/quick_log$L.err; echo "log$L rc=$?"
  cat gpurun_out/quick_log$L.json | cut -c1-120; tail -n 3 gpurun_out/quick_log$L.err
done
