# Single-GPU pass: parity tests, bench line, optional extra commands in $EXTRA
set -x
mkdir -p gpurun_out
TAG=${TAG:-exp}
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/gpu_tests_$TAG.log
python bench.py --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_$TAG.json')); print(d['ms_per_step'], d['e2e']['ms_per_step']); print(d['phase_ms_per_step'])"
tail -n 3 gpurun_out/bench_$TAG.err
for L in ${SIZES:-}; do
  python bench.py --log-rows $L --quick --steps 5 --warmup 1 --no-cpu-baseline 2>/dev/null | cut -c1-60
done
if [ -n "${TRACE:-}" ]; then
  python bench.py --trace gpurun_out/trace_$TAG.json --no-cpu-baseline > gpurun_out/trace.log 2>&1
  python tools/trace_gaps.py gpurun_out/trace_$TAG.json 14 > gpurun_out/trace_gaps_$TAG.txt 2>&1
  head -18 gpurun_out/trace_gaps_$TAG.txt; gzip -f gpurun_out/trace_$TAG.json
fi
if [ -n "${SMOKE:-}" ]; then python -c "import __graft_entry__ as g; g.smoke()"; fi
