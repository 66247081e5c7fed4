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
