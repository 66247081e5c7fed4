# periodic columns / auxiliary-segment AIR tests + a short bench line with the re-measured ALU peak
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_y1}
export AERO_B200_NO_BUILD=1
timeout 600 python -m pytest tests/test_air_fib2.py -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/${TAG}_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-lde-download --steps 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['peak'], d['roofline']['ntt']['int_frac'])"
