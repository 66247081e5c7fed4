set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_y3}
export AERO_B200_NO_BUILD=1
timeout 240 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "air_program or aux_builder or rejects" > gpurun_out/${TAG}_tests_sharded_air.log 2>&1; echo "sharded air rc=$?"
tail -25 gpurun_out/${TAG}_tests_sharded_air.log
timeout 240 python -m pytest tests/test_air_fib2.py tests/test_gpu_parity.py -m gpu -x -q -k "constraint or air or fib2 or periodic or bitwise or aux_segment or prove" > gpurun_out/${TAG}_tests_single.log 2>&1; echo "single rc=$?"
tail -5 gpurun_out/${TAG}_tests_single.log
