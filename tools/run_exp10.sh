# GPU pass: parity tests (early-hash batches, copy pool behind the pageable staging), bench incl. e2e_pageable, edge-batch variant
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_x10}
export AERO_B200_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/${TAG}_tests.log
for e in -2 2; do
  timeout 300 python bench.py --no-cpu-baseline --no-lde-download --steps 10 --upload-edge-cols $e > gpurun_out/${TAG}_bench_edge$e.json 2> gpurun_out/${TAG}_bench_edge$e.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench_edge$e.json'))
print('edge=$e', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'pageable', round(d['e2e_pageable']['ms_per_step'],2))
PY
done
