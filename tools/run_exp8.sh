# GPU pass: parity tests (AIR evaluator with liveness slots, mulfib2, random programs) + evaluator throughput
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_x8}
export AERO_B200_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/${TAG}_tests.log
for nodes in 128 512 2048; do
  timeout 300 python tools/air_bench.py --nodes $nodes 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_air_bench.jsonl
done
timeout 300 python tools/air_bench.py --nodes 512 --ce-blowup 2 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_air_bench.jsonl
