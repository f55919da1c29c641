set -x
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/multigpu_check.py 2>&1 | grep -v "^W\|^\[W" | tail -8
for X in ${EXCH:-p2p nccl}; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 30 --warmup 5 --exchange $X > gpurun_out/bench_mg${N}_$X.json 2> gpurun_out/bench_mg${N}_$X.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_mg${N}_$X.json').read().strip().splitlines()[-1])
print('N=$N $X', d['value'], 'Mrays/s', d['ms_per_step'], 'ms/frame  e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])
"; grep -v "^W\|^\[W" gpurun_out/bench_mg${N}_$X.err | grep -iE "error|Traceback" | head -5
done
# frame-sharded partition (opt-in, weak scaling): every rank renders whole frames of its own, nothing exchanged
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 5 --partition frames > gpurun_out/bench_${N}gpu_frames.json 2> gpurun_out/bench_${N}gpu_frames.err; tail -c 600 gpurun_out/bench_${N}gpu_frames.json; tail -2 gpurun_out/bench_${N}gpu_frames.err
