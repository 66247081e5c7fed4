# Round-end two-GPU pass (gpurun --gpus 2): IPC window parity test, ONE 2^20-row proof over two GPUs (the driver's
# --gpus 2 shape, with the in-run parity checks), one 2^22-row proof (three-pass NTT plans, coset-sharded)
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_final}
export AERO_B200_NO_BUILD=1
timeout 600 python -m pytest tests/test_gpu_window.py -m gpu -x -q > gpurun_out/${TAG}_tests_2gpu.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${TAG}_tests_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-lde-download > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "n2 rc=$?"
timeout 900 $TR bench.py --gpus 2 --log-rows 22 --steps 3 --warmup 1 --no-lde-download > gpurun_out/${TAG}_bench_n2_log22.json 2> gpurun_out/${TAG}_bench_n2_log22.err; echo "log22 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_final_bench_n2*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('parity'), d.get('independent_proofs',{}).get('ms_per_step')); print(d['phase_ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
for f in gpurun_out/${TAG}_bench_n2*.err; do tail -n 3 $f; done
