# Three-pass NTT plans (n = 2^10 * 2^10 * n0 above 2^20 points) on ONE B200: parity tests, the headline
# bench, and 2^22 / 2^24-row commitments and proofs with the split on (default) and off (AERO_NTT_OUTER=0).
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_ntt3}
export AERO_B200_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/${TAG}_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-lde-download --steps 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step']); print(d.get("phase_ms_per_step"))"
for o in -1 0; do
  AERO_NTT_OUTER=$o timeout 600 python tools/sweep.py --sizes 21,22,23,24 --widths 8,72 --reps 2 --out gpurun_out/${TAG}_sweep_outer$o.jsonl > /dev/null 2> gpurun_out/${TAG}_sweep_outer$o.err
  python - <<PY
import json
for l in open('gpurun_out/${TAG}_sweep_outer$o.jsonl'):
    d=json.loads(l); print('outer=$o', d['log_rows'], d['cols'], 'commit %.2f interp %.2f lde %.2f hash %.2f bfly/s %.3g' % (d['commit_ms'], d['interpolate_ms'], d['lde_ms'], d['hash_rows_ms'], d['ntt_bfly_s']))
PY
  for L in 22 24; do
    AERO_NTT_OUTER=$o timeout 600 python bench.py --log-rows $L --quick --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_log${L}_outer$o.json 2>/dev/null
    python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_log${L}_outer$o.json')); print('outer=$o log_rows=$L', d['ms_per_step']); print(d.get("phase_ms_per_step"))"
  done
done
