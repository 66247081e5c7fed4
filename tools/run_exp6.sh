# ncu source-level capture of the LDE NTT passes (per-line instruction counts and stall samples)
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_x6}
export AERO_B200_NO_BUILD=1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:dft_pass' -s 2 -c 2 -o gpurun_out/${TAG}_ncu_src -f python bench.py --quick --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_src.log 2>&1
ncu -i gpurun_out/${TAG}_ncu_src.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_ncu_src_sass.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_ncu_src.ncu-rep --page source --csv --print-source cuda > gpurun_out/${TAG}_ncu_src_cuda.csv 2>/dev/null
ls -la gpurun_out/${TAG}_ncu_src*
rm -f gpurun_out/${TAG}_ncu_src.ncu-rep
gzip -f gpurun_out/${TAG}_ncu_src_sass.csv
head -c 1500 gpurun_out/${TAG}_ncu_src_cuda.csv
