set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_y7}
export AERO_B200_NO_BUILD=1
timeout 200 python -m pytest tests/test_air_fib2.py -m gpu -x -q > gpurun_out/${TAG}_tests_air.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests_air.log
for b in -1 0 4 8; do
  timeout 60 python tools/air_bench.py --air bitwise --log-rows 20 --reps 3 --blocks-per-sm $b 2>/dev/null | cut -c1-330 >> gpurun_out/${TAG}_air_bench.jsonl
done
for b in -1 0; do
  timeout 60 python tools/air_bench.py --nodes 512 --log-rows 20 --reps 2 --blocks-per-sm $b 2>/dev/null | cut -c1-300 >> gpurun_out/${TAG}_air_bench.jsonl
done
cat gpurun_out/${TAG}_air_bench.jsonl
