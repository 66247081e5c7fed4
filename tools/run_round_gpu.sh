# One round-end GPU pass: parity tests, bench (both arms), ncu launch list, ncu --set full of the hot
# kernels (converted to CSV on the box: gpurun_out/ is capped at 64 MiB), CUPTI timeline of one step.
set -x
mkdir -p gpurun_out
TAG=${TAG:-v10}
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/gpu_tests_$TAG.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
cat gpurun_out/bench_${TAG}_ref.json
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --quick --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:dft_pass' -c 4 -o gpurun_out/ncu_${TAG}_ntt -f python bench.py --quick --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_ntt.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:hash_rows|merkle_subtree' -c 2 -o gpurun_out/ncu_${TAG}_hash -f python bench.py --quick --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_hash.log 2>&1
for r in ntt hash; do
  ncu -i gpurun_out/ncu_${TAG}_$r.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_${r}_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_${TAG}_$r.ncu-rep --page source --csv > gpurun_out/ncu_${TAG}_${r}_source.csv 2>/dev/null
done
python bench.py --trace gpurun_out/trace_$TAG.json --no-cpu-baseline > gpurun_out/trace.log 2>&1
python tools/trace_gaps.py gpurun_out/trace_$TAG.json > gpurun_out/trace_gaps_$TAG.txt 2>&1
gzip -f gpurun_out/trace_$TAG.json
rm -f gpurun_out/ncu_${TAG}_ntt.ncu-rep
du -sm gpurun_out; ls -la gpurun_out
# keep the merge under the cap: drop the largest .ncu-rep files first
for f in $(ls -S gpurun_out/*.ncu-rep); do
  [ $(du -sm gpurun_out | cut -f1) -lt 60 ] && break
  rm -f $f
done
du -sm gpurun_out
