set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_y6}
export AERO_B200_NO_BUILD=1
timeout 240 ncu --set full --clock-control none --import-source on -k 'regex:air_evaluate' -c 2 -o gpurun_out/${TAG}_ncu_air -f python tools/air_bench.py --air bitwise --log-rows 20 --reps 1 > gpurun_out/ncu_air.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}_ncu_air.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_air_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_ncu_air_raw.csv > gpurun_out/${TAG}_ncu_air.txt 2>&1
rm -f gpurun_out/${TAG}_ncu_air.ncu-rep
tail -40 gpurun_out/${TAG}_ncu_air.txt
