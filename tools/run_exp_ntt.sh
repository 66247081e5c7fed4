# NTT experiment pass on ONE B200: parity tests of the TMA (BULK) tile loads, A/B of the bulk loads and of the
# column-fastest pass-1 grid, ncu --set full of the NTT passes.  TAG names the outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_tma}
export AERO_B200_NO_BUILD=1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/${TAG}_tests.log
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  AERO_NTT_BULK=$1 AERO_NTT_COLFAST=$2 timeout 300 python bench.py --no-cpu-baseline --no-lde-download --steps 10 > gpurun_out/${TAG}_bench_b$1_c$2.json 2> gpurun_out/${TAG}_bench_b$1_c$2.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench_b$1_c$2.json'))
p=d['phase_ms_per_step']
print('bulk=$1 colfast=$2', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k:p[k] for k in ('interpolate_w72','lde_w72','lde_w9','lde_w8','hash_rows_w72')})
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:dft_pass' -c 4 -o gpurun_out/${TAG}_ncu_ntt -f python bench.py --quick --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_ntt.log 2>&1
ncu -i gpurun_out/${TAG}_ncu_ntt.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_ntt_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_ncu_ntt_raw.csv > gpurun_out/${TAG}_ncu_ntt.txt 2>&1
rm -f gpurun_out/${TAG}_ncu_ntt.ncu-rep
grep -E "Kernel Name|gpu__time_duration|dram__bytes|pipe_alu.avg|issue_active" gpurun_out/${TAG}_ncu_ntt.txt | cut -c1-150
