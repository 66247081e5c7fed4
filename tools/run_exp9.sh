set -x
mkdir -p gpurun_out
export AERO_B200_NO_BUILD=1
for ce in 2 4 8; do timeout 300 python tools/air_bench.py --nodes 512 --ce-blowup $ce 2>&1 | tail -1 | tee -a gpurun_out/r02_x9_air_bench.jsonl; done
timeout 300 python tools/air_bench.py --nodes 128 --ce-blowup 8 2>&1 | tail -1 | tee -a gpurun_out/r02_x9_air_bench.jsonl
timeout 300 python tools/air_bench.py --nodes 2048 --ce-blowup 8 2>&1 | tail -1 | tee -a gpurun_out/r02_x9_air_bench.jsonl
