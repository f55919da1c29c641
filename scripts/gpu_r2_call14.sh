# Round 2, call 14 (8 GPUs; call 6 again with the round's final library): the N-GPU paths as BASELINE names them -- the 4K frame split over 8 / 4 GPUs (fused frame kernel for
# small shares), configs[3] C4 (1024^3 terrain) and configs[4] C5 (an edit per frame, device-side commands) on 8 GPUs, and the
# multi-GPU parity check (frames / edits / fingerprints vs the oracle) as pytest
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
export VXRT_MULTIGPU_LOG=$PWD/gpurun_out/r2c14_multigpu.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "8" > gpurun_out/r2c14_pytest_multi.log 2>&1; tail -3 gpurun_out/r2c14_pytest_multi.log
run() {  # run N name args...
  N=$1; name=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 297$N$N bench.py --gpus $N --steps 30 --warmup 5 "$@" > gpurun_out/r2c14_bench_$name.json 2> gpurun_out/r2c14_bench_$name.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r2c14_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', d.get('ms_per_step_spread',{}).get('median'), d['roofline'].get('frame_kernel',{}).get('ms'), d['value'], d['unit'], d['ms_per_step'], 'ms  e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d.get('parity',{}).get('mismatched_pixels'), d.get('parity',{}).get('frame_fnv'), d['clocks']['sm_mhz'])
except Exception as e:
    print('$name FAILED', e)
P
  grep -v "^W\|^\[W" gpurun_out/r2c14_bench_$name.err | grep -iE "error|Traceback|assert" | head -5
}
run 8 8gpu
run 4 4gpu
run 8 8gpu_C4 --workload C4_terrain_4k
run 8 8gpu_C5 --workload C5_edits_4k
run 2 2gpu
run 8 8gpu_hoststores --e2e-host-stores
ls -la gpurun_out | tail -14
