set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_y4}
export AERO_B200_NO_BUILD=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench_n2.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_n2.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['scaling'], d['parity'])"
timeout 200 python -m pytest tests/test_gpu_window.py -m gpu -x -q 2>&1 | tail -3
