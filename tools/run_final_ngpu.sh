# Round-end multi-GPU pass (gpurun --gpus $G): ONE 2^20-row proof over G GPUs (the driver's --gpus G shape, with the
# in-run parity checks) and one 2^24-row proof (BASELINE configs[3]; three-pass NTT plans, coset-sharded)
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_final}
G=${G:-8}
export AERO_B200_NO_BUILD=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus $G --steps 10 --warmup 3 --no-lde-download > gpurun_out/${TAG}_bench_n$G.json 2> gpurun_out/${TAG}_bench_n$G.err; echo "n$G rc=$?"
if [ -n "${BIG:-}" ]; then
timeout 900 $TR bench.py --gpus $G --log-rows $BIG --quick --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_n${G}_log${BIG}_quick.json 2> gpurun_out/${TAG}_bench_n${G}_log${BIG}.err; echo "log$BIG rc=$?"
fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_final_bench_n[48]*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, d['ms_per_step'], d.get('value'), (d.get('e2e') or {}).get('ms_per_step'), d.get('parity')); print(d.get('phase_ms_per_step') or d.get('phase_ms'))
    except Exception as e: print(f, 'ERR', e)
PY
for f in gpurun_out/${TAG}_bench_n$G*.err; do tail -n 3 $f | cut -c1-300; done
