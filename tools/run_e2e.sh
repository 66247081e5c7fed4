# diagnostics: run-to-run spread of the bench line (device-resident step and host-buffer step)
mkdir -p gpurun_out
for i in 1 2 3; do
python bench.py --no-cpu-baseline > gpurun_out/bench_rep$i.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_rep$i.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])"
done
