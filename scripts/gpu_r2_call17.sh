# Round 2, call 17 (2 GPUs): the whole peer-memory chain (begin -> render -> signal -> owner's wait -> release) on programmatic
# dependent launches, no event records inside it (chain span from the device clock)
set -x
mkdir -p gpurun_out
export VXRT_MULTIGPU_LOG=$PWD/gpurun_out/r2c17_multigpu.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c17_pytest.log 2>&1; tail -5 gpurun_out/r2c17_pytest.log
run() {  # run name env... -- args...
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --steps 30 --warmup 5 "$@" > gpurun_out/r2c17_bench_$name.json 2> gpurun_out/r2c17_bench_$name.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r2c17_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'], 'ms  e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('frame_check'), d.get('parity',{}).get('mismatched_pixels'), d.get('parity',{}).get('frame_fnv'), d.get('per_rank'))
except Exception as e:
    print('$name FAILED', e)
P
  grep -v "^W\|^\[W" gpurun_out/r2c17_bench_$name.err | grep -iE "error|Traceback|assert|timed out" | head -5
}
run 2gpu VXRT_P2P_PDL=1 --
run 2gpu_fused VXRT_FUSION=1 --
run 2gpu_fused_nopdl VXRT_FUSION=1 VXRT_P2P_PDL=0 --
run 2gpu_C5_host VXRT_FUSION=1 -- --workload C5_edits_4k --e2e-path host
