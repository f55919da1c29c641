# Round 2, call 8 (2 GPUs): call 7 again after the fix (the back-pressure wait overlaps the first render kernel only on importers whose
# GPU is not the owner's: on the owner's device the release waited behind the render kernel's undispatched blocks)
set -x
mkdir -p gpurun_out
export VXRT_MULTIGPU_LOG=$PWD/gpurun_out/r2c8_multigpu.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c8_pytest.log 2>&1; tail -5 gpurun_out/r2c8_pytest.log
run() {  # run name env... -- args...
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --steps 30 --warmup 5 "$@" > gpurun_out/r2c8_bench_$name.json 2> gpurun_out/r2c8_bench_$name.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r2c8_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'], 'ms  e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('frame_check'), d.get('parity',{}).get('mismatched_pixels'), d.get('parity',{}).get('frame_fnv'), d['roofline']['kernel'], d['roofline'].get('frame_kernel',{}).get('ms'), d['roofline']['kernels']['primary_kernel']['ms'], d['roofline']['kernels']['shade_kernel']['ms'])
except Exception as e:
    print('$name FAILED', e)
P
  grep -v "^W\|^\[W" gpurun_out/r2c8_bench_$name.err | grep -iE "error|Traceback|assert|timed out" | head -5
}
run 2gpu_pdl VXRT_P2P_PDL=1 --
run 2gpu_fused_pdl VXRT_FUSION=1 --
run 2gpu_host VXRT_P2P_PDL=1 -- --e2e-path host
ls -la gpurun_out | tail -8
