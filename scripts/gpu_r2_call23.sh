# Round 2, call 23 (8 GPUs): the 4K frame on 8 GPUs with the round's final library (one run)
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29788 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r2c23_bench_8gpu.json 2> gpurun_out/r2c23_bench_8gpu.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2c23_bench_8gpu.json').read().strip().splitlines()[-1])
print('8gpu', d['value'], d['ms_per_step'], 'ms  e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('frame_check'), d.get('parity',{}).get('mismatched_pixels'), d.get('parity',{}).get('frame_fnv'), d.get('per_rank'))
P
grep -v "^W\|^\[W" gpurun_out/r2c23_bench_8gpu.err | grep -iE "error|Traceback|assert|timed out" | head -5
