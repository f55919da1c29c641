# Round 2, call 7 (2 GPUs): flag protocol of the peer-memory frames chained with programmatic dependent launches, tile-row partition
# + strip DMA into the shared host frame
set -x
mkdir -p gpurun_out
export VXRT_MULTIGPU_LOG=$PWD/gpurun_out/r2c7_multigpu.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c7_pytest.log 2>&1; tail -5 gpurun_out/r2c7_pytest.log
run() {  # run name env... -- args...
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --steps 30 --warmup 5 "$@" > gpurun_out/r2c7_bench_$name.json 2> gpurun_out/r2c7_bench_$name.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r2c7_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'], 'ms  e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('frame_check'), d.get('parity',{}).get('mismatched_pixels'), d.get('parity',{}).get('frame_fnv'), d['roofline']['kernel'], d['roofline'].get('frame_kernel',{}).get('ms'), d['roofline']['kernels']['primary_kernel']['ms'], d['roofline']['kernels']['shade_kernel']['ms'])
except Exception as e:
    print('$name FAILED', e)
P
  grep -v "^W\|^\[W" gpurun_out/r2c7_bench_$name.err | grep -iE "error|Traceback|assert|timed out" | head -5
}
run 2gpu_fused_pdl VXRT_FUSION=1 -- --e2e-path host
run 2gpu_fused_nopdl VXRT_FUSION=1 VXRT_P2P_PDL=0 -- --e2e-path host --e2e-host-stores
run 2gpu_pdl VXRT_P2P_PDL=1 --
run 2gpu_nopdl VXRT_P2P_PDL=0 --
ls -la gpurun_out | tail -12
