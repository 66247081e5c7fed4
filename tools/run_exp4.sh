# GPU pass: parity tests (incl. the device AIR evaluator), bench line, CUPTI timeline of the host-buffer step
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_x4}
export AERO_B200_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/${TAG}_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-lde-download --steps 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('e2e_pageable',{}).get('ms_per_step')); print(d['phase_ms_per_step'])"
timeout 300 python bench.py --trace gpurun_out/${TAG}_trace_host.json --trace-host --no-cpu-baseline > gpurun_out/trace.log 2>&1
python tools/trace_gaps.py gpurun_out/${TAG}_trace_host.json 25 kernels > gpurun_out/${TAG}_trace_gaps_host_kernels.txt 2>&1
python tools/trace_gaps.py gpurun_out/${TAG}_trace_host.json 10 > gpurun_out/${TAG}_trace_gaps_host_all.txt 2>&1
head -30 gpurun_out/${TAG}_trace_gaps_host_kernels.txt; head -6 gpurun_out/${TAG}_trace_gaps_host_all.txt
python - <<'PY'
import json
ev=json.load(open('gpurun_out/r02_x4_trace_host.json'))["traceEvents"]
k=[e for e in ev if e.get("ph")=="X" and e.get("cat") in ("kernel","gpu_memcpy")]
t0=min(e["ts"] for e in k)
# coarse timeline: per 1 ms bucket, busy us of kernels and of copies
import collections
bk=collections.defaultdict(float); bc=collections.defaultdict(float)
for e in k:
    b=int((e["ts"]-t0)//1000)
    (bk if e["cat"]=="kernel" else bc)[b]+=e["dur"]
for b in range(0, int(max(list(bk)+list(bc)))+1):
    print(b, round(bk[b]), round(bc[b]))
PY
gzip -f gpurun_out/${TAG}_trace_host.json
