set -x
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/multigpu_check.py 2>&1 | grep -v "^W\|^\[W" | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_mg$N.json 2> gpurun_out/bench_mg$N.err; tail -c 2500 gpurun_out/bench_mg$N.json; grep -v "^W\|^\[W" gpurun_out/bench_mg$N.err | tail -8
