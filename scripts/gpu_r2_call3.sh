# Round 2, third GPU call (2 GPUs): overlapped passes (programmatic dependent launch), device-side edit commands, the N-GPU paths.
set -x
mkdir -p gpurun_out
export VXRT_MULTIGPU_LOG=$PWD/gpurun_out/r2c3_multigpu.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c3_pytest.log 2>&1; tail -6 gpurun_out/r2c3_pytest.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r2c3_bench_1gpu.json 2> gpurun_out/r2c3_bench_1gpu.err; tail -c 1500 gpurun_out/r2c3_bench_1gpu.json; tail -3 gpurun_out/r2c3_bench_1gpu.err
for v in 0 1; do VXRT_OVERLAP=$v python scripts/exp_probe.py | tee -a gpurun_out/r2c3_probe_overlap.jsonl | cut -c1-600; done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711"
$TR bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2c3_bench_2gpu.json 2> gpurun_out/r2c3_bench_2gpu.err; tail -c 1200 gpurun_out/r2c3_bench_2gpu.json; tail -3 gpurun_out/r2c3_bench_2gpu.err
$TR bench.py --gpus 2 --steps 30 --warmup 5 --workload C5_edits_4k > gpurun_out/r2c3_bench_2gpu_C5.json 2> gpurun_out/r2c3_bench_2gpu_C5.err; tail -c 1200 gpurun_out/r2c3_bench_2gpu_C5.json; tail -3 gpurun_out/r2c3_bench_2gpu_C5.err
python bench.py --steps 20 --warmup 5 --workload C5_edits_4k > gpurun_out/r2c3_bench_1gpu_C5.json 2> gpurun_out/r2c3_bench_1gpu_C5.err; tail -c 800 gpurun_out/r2c3_bench_1gpu_C5.json; tail -3 gpurun_out/r2c3_bench_1gpu_C5.err
ls -la gpurun_out | tail -12
