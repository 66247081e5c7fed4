# Eight-GPU pass (gpurun --gpus 8): independent proofs per rank, one 2^20-row proof sharded by LDE coset,
# one 2^24-row proof sharded by LDE coset (BASELINE config 4).  Every leg runs under its own timeout.
set -x
mkdir -p gpurun_out
G=${G:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29551"
[ -n "${WEAK:-}" ] && timeout 300 $TR bench.py --gpus $G --no-cpu-baseline > gpurun_out/bench_n${G}_weak.json 2> gpurun_out/bench_n${G}_weak.err; echo "weak rc=$?"
timeout 300 $TR bench.py --gpus $G --shard-proof --no-cpu-baseline > gpurun_out/bench_n${G}_shard_window.json 2> gpurun_out/bench_n${G}_shard_window.err; echo "shard rc=$?"
[ -n "${MID:-}" ] && timeout 300 $TR bench.py --gpus $G --shard-proof --log-rows 22 --quick --steps 3 --no-cpu-baseline > gpurun_out/bench_n${G}_shard_log22.json 2> gpurun_out/bench_n${G}_shard_log22.err; echo "log22 rc=$?"
if [ -n "${BIG:-}" ]; then
timeout 480 $TR bench.py --gpus $G --shard-proof --log-rows 24 --quick --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_n${G}_shard_log24.json 2> gpurun_out/bench_n${G}_shard_log24.err; echo "log24 rc=$?"
fi
for f in gpurun_out/bench_n${G}_*.json; do echo $f; grep '^{' $f | tail -n 1 | cut -c1-160; done
for f in gpurun_out/*.err; do echo $f; tail -n 2 $f | cut -c1-300; done
