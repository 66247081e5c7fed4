# GPU pass: parity tests, A/B of the fused FRI fold-and-hash and of early chained hashing on the host-buffer path
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_x5}
export AERO_B200_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/${TAG}_tests.log
for cfg in "1 2" "0 2" "1 0" "1 1" "1 3" "1 2"; do
  set -- $cfg
  AERO_FRI_FUSED=$1 AERO_HASH_EARLY=$2 timeout 300 python bench.py --no-cpu-baseline --no-lde-download --steps 10 > gpurun_out/${TAG}_bench_f$1_h$2.json 2> gpurun_out/${TAG}_bench_f$1_h$2.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench_f$1_h$2.json'))
p=d['phase_ms_per_step']
print('fri_fused=$1 hash_early=$2', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'pageable', round(d.get('e2e_pageable',{}).get('ms_per_step',0),2), {k:p[k] for k in ('fri_commit','fri_fold','hash_rows_w72')})
PY
done
