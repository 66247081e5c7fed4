# Two-GPU pass (gpurun --gpus 2): window parity test, weak-scaling bench, one sharded proof (window and NCCL routes)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_window.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/gpu_tests_2gpu.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/gpu_tests_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_n2_weak.json 2> gpurun_out/bench_n2_weak.err; echo "weak rc=$?"
timeout 600 $TR bench.py --gpus 2 --shard-proof --no-cpu-baseline > gpurun_out/bench_n2_shard_window.json 2> gpurun_out/bench_n2_shard_window.err; echo "shard rc=$?"
timeout 600 $TR bench.py --gpus 2 --shard-proof --nccl-exchange --no-cpu-baseline > gpurun_out/bench_n2_shard_nccl.json 2> gpurun_out/bench_n2_shard_nccl.err; echo "nccl rc=$?"
timeout 900 $TR bench.py --gpus 2 --shard-proof --log-rows 22 --steps 3 --no-cpu-baseline > gpurun_out/bench_n2_shard_log22.json 2> gpurun_out/bench_n2_shard_log22.err; echo "log22 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_n2_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
for f in gpurun_out/*.err; do tail -n 3 $f; done
