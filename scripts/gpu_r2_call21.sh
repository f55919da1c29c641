# Round 2, call 21 (2 GPUs): strip path of the host frames waits for the display rank's release only before it queues the copy
set -x
mkdir -p gpurun_out
export VXRT_MULTIGPU_LOG=$PWD/gpurun_out/r2c21_multigpu.log
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -k "multigpu or host_frames or partition or pipelined" > gpurun_out/r2c21_pytest.log 2>&1; tail -4 gpurun_out/r2c21_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --steps 30 --warmup 5 --e2e-path host > gpurun_out/r2c21_bench_2gpu_host.json 2> gpurun_out/r2c21_bench_2gpu_host.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2c21_bench_2gpu_host.json').read().strip().splitlines()[-1])
print('2gpu_host', d['value'], d['ms_per_step'], 'ms  e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('frame_check'), d.get('parity',{}).get('mismatched_pixels'), d.get('parity',{}).get('frame_fnv'))
P
grep -v "^W\|^\[W" gpurun_out/r2c21_bench_2gpu_host.err | grep -iE "error|Traceback|assert|timed out" | head -5
VXRT_FUSION=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --steps 30 --warmup 5 --e2e-path host --workload C2_1080p > gpurun_out/r2c21_bench_2gpu_host_1080p.json 2> gpurun_out/r2c21_bench_2gpu_host_1080p.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2c21_bench_2gpu_host_1080p.json').read().strip().splitlines()[-1])
print('2gpu_host_1080p', d['value'], d['ms_per_step'], 'ms  e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('frame_check'), d.get('parity',{}).get('mismatched_pixels'))
P
