set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_v6.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/gpu_tests_v6.log
python bench.py --no-cpu-baseline > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; echo "bench rc=$?"
cat gpurun_out/bench_v6.json; tail -3 gpurun_out/bench_v6.err
python bench.py --no-cpu-baseline --upload-batch-cols 4 > gpurun_out/bench_v6_ub4.json 2> gpurun_out/bench_v6_ub4.err
python -c "
import json
for f in ['bench_v6','bench_v6_ub4']:
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['ms_per_step'], d['e2e']['ms_per_step'])
"
AERO_NTT_TILE=4 python bench.py --quick --steps 5 --no-cpu-baseline > gpurun_out/quick_tile4.json 2>&1
cat gpurun_out/quick_tile4.json
python bench.py --quick --steps 5 --no-cpu-baseline > gpurun_out/quick_tile8.json 2>&1
cat gpurun_out/quick_tile8.json
python bench.py --trace gpurun_out/trace_v6.json --no-cpu-baseline > gpurun_out/trace.log 2>&1
python tools/trace_gaps.py gpurun_out/trace_v6.json 12 > gpurun_out/trace_gaps_v6.txt 2>&1
head -16 gpurun_out/trace_gaps_v6.txt
gzip -f gpurun_out/trace_v6.json
