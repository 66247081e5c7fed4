# Final evidence of the round on ONE B200: all parity tests, smoke(), the bench line, the reference arm (short budget),
# the device AIR evaluator on Miden's bitwise chiplet, memcheck over the AIR evaluator cases.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_last}
export AERO_B200_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/${TAG}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_pageable']['ms_per_step'], d['roofline']['frac'], d['roofline']['ntt']['int_frac'], d['cpu_baseline'], d['clocks'])"
timeout 200 python bench.py --impl reference --steps 1 --warmup 1 --ref-budget-s 25 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
timeout 200 python tools/air_bench.py --air bitwise --log-rows 20 --prove > gpurun_out/${TAG}_air_bench_bitwise.jsonl 2> gpurun_out/${TAG}_air_bench.err; echo "air rc=$?"; cat gpurun_out/${TAG}_air_bench_bitwise.jsonl
timeout 150 compute-sanitizer --tool memcheck --log-file gpurun_out/${TAG}_sanitizer_memcheck_air.log python -m pytest tests/test_air_fib2.py -m gpu -x -q -k "(bitwise or periodic or aux_segment or mulfib2) and canonical and not 14 and not 12 and not 13" > gpurun_out/${TAG}_san_air.out 2>&1; echo "san rc=$?"; tail -2 gpurun_out/${TAG}_san_air.out; tail -2 gpurun_out/${TAG}_sanitizer_memcheck_air.log
