# GPU pass: parity tests (new: sharded three-pass, alignment fallback, FRI modes), bench, host timing marks
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_x7}
export AERO_B200_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/${TAG}_tests.log
for f in 1 0; do
  AERO_FRI_FUSED=$f timeout 300 python bench.py --no-cpu-baseline --no-lde-download --steps 10 > gpurun_out/${TAG}_bench_f$f.json 2> gpurun_out/${TAG}_bench_f$f.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench_f$f.json'))
p=d['phase_ms_per_step']
print('fri_fused=$f', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {k:p[k] for k in ('fri_commit','fri_fold','lde_w72','interpolate_w72','hash_rows_w72')})
PY
done
AERO_HOST_TIMING=1 timeout 300 python bench.py --quick --steps 3 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_host_marks.txt
tail -40 gpurun_out/${TAG}_host_marks.txt
