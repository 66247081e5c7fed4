set -x
mkdir -p gpurun_out
for L in 16 18 22 23; do
  timeout 600 python bench.py --log-rows $L --quick --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/quick_log$L.json 2> gpurun_out/quick_log$L.err
  cat gpurun_out/quick_log$L.json | cut -c1-200
done
timeout 900 python tools/sweep.py --logs 16,18,20,22,24 --cols 1,8,72,255 --reps 3 --out gpurun_out/sweep_v7.jsonl > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"
tail -n 4 gpurun_out/sweep.log | cut -c1-400
