# host-frame path: one-GPU protocol test, then (on an N-GPU box) the cross-process check and the bench with --e2e-path host
set -x
mkdir -p gpurun_out
N=${1:-1}
python -m pytest tests/test_gpu_parity.py -x -q -k "host_frames or peer_memory or banded or pipelined" 2>&1 | tail -15
if [ "$N" -gt 1 ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/multigpu_check.py 2>&1 | grep -v "^W\|^\[W" | tail -9
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 30 --warmup 5 --no-extra --e2e-path host > gpurun_out/bench_mg${N}_host.json 2> gpurun_out/bench_mg${N}_host.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_mg${N}_host.json').read().strip().splitlines()[-1])
print('N=$N host', d['value'], 'Mrays/s', d['ms_per_step'], 'ms/frame  e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('frame_check'), d['e2e']['api'][:60])
"; grep -v "^W\|^\[W" gpurun_out/bench_mg${N}_host.err | grep -iE "error|Traceback|unavailable" | head -5
fi
