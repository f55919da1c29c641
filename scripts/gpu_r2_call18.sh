# Round 2, call 18 (8 GPUs; the closing multi-GPU run): the 4K frame on 8 GPUs with wide blocks (+ per-rank figures), C5 with the pipelined e2e path
set -x
mkdir -p gpurun_out
run() {  # run N name args...
  N=$1; name=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 297$N$N bench.py --gpus $N --steps 30 --warmup 5 "$@" > gpurun_out/r2c18_bench_$name.json 2> gpurun_out/r2c18_bench_$name.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r2c18_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'], 'ms  e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d.get('parity',{}).get('mismatched_pixels'), d.get('parity',{}).get('frame_fnv'), d.get('per_rank'))
except Exception as e:
    print('$name FAILED', e)
P
  grep -v "^W\|^\[W" gpurun_out/r2c18_bench_$name.err | grep -iE "error|Traceback|assert" | head -5
}
run 8 8gpu
run 8 8gpu_C4 --workload C4_terrain_4k
run 4 4gpu
