set -x
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader
nvidia-smi topo -m 2>/dev/null | head -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/multigpu_check.py 2>&1 | grep -v "^W\|^\[W" | tail -14
for X in p2p nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 30 --warmup 5 --exchange $X > gpurun_out/bench_mg${N}_$X.json 2> gpurun_out/bench_mg${N}_$X.err; tail -c 1800 gpurun_out/bench_mg${N}_$X.json; grep -v "^W\|^\[W" gpurun_out/bench_mg${N}_$X.err | tail -6
done
