# Round-end evidence on ONE B200: parity tests, smoke(), the bench line (with the CPU baseline and the separately
# timed LDE download), the reference arm (short budget), ncu launch list + ncu --set full of the hot kernels,
# compute-sanitizer memcheck / racecheck over smoke() and over a two-rank sharded proof.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_final}
export AERO_B200_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['ntt']['int_frac'], d['cpu_baseline']); print(d['phase_ms_per_step'])"
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 --ref-budget-s 30 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --quick --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:hash_rows_kernel|merkle_subtree' -c 3 -o gpurun_out/${TAG}_ncu_hash -f python bench.py --quick --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_hash.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:dft_pass' -c 4 -o gpurun_out/${TAG}_ncu_ntt -f python bench.py --quick --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_ntt.log 2>&1
for r in hash ntt; do
  ncu -i gpurun_out/${TAG}_ncu_$r.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_${r}_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/${TAG}_ncu_${r}_raw.csv > gpurun_out/${TAG}_ncu_${r}.txt 2>&1
  rm -f gpurun_out/${TAG}_ncu_$r.ncu-rep
done
cat > /tmp/san_sharded.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import aero_b200
from aero_b200 import make_divisor
from aero_b200.sharded import window_bytes
from oracle import stark_oracle as so
P = so.P
logn, wm, wa, world = 8, 6, 3, 2
n = 1 << logn
main, aux, ce = so.synthetic_trace(wm, n), so.synthetic_trace(wa, n, 0xAE210000), so.synthetic_trace(2, 8 * n, 0xCE)
divs = [so.Divisor(n, 1, [pow(so.root_of_unity(logn), n - 1, P)]), so.Divisor(1, 1, [])]
ref = so.prove(main, aux, ce, divs, b"san")
g = aero_b200.Group([0] * world, window_bytes(logn, wm + wa, world), form=aero_b200.AERO_FORM_CANONICAL)
g.set_option("force_host_sync", 1)   # the sanitizer serialises kernels: device-side flag barriers cannot progress
for it in range(2):
    got = g.prove(main, aux, ce, [make_divisor(d.a, d.b, d.exemptions) for d in divs], b"san")
    assert got == ref.proof_bytes
g.close()
print("sharded proof under the sanitizer: identical to the oracle's")
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/${TAG}_sanitizer_${tool}_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_${tool}_smoke.out 2>&1; tail -2 gpurun_out/san_${tool}_smoke.out
  CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/${TAG}_sanitizer_${tool}_sharded.log python /tmp/san_sharded.py > gpurun_out/san_${tool}_sharded.out 2>&1; tail -2 gpurun_out/san_${tool}_sharded.out
  tail -3 gpurun_out/${TAG}_sanitizer_${tool}_smoke.log gpurun_out/${TAG}_sanitizer_${tool}_sharded.log
done
du -sm gpurun_out
