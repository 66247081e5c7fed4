set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_y2}
export AERO_B200_NO_BUILD=1
timeout 600 python -m pytest tests/test_air_fib2.py -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/${TAG}_tests.log
