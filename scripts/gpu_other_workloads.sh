set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_host.py -x -q 2>&1 | tail -4
python bench.py --workload C4_terrain_4k --steps 20 --warmup 3 --no-extra > gpurun_out/bench_c4_v6.json 2> gpurun_out/bench_c4_v6.err; tail -c 600 gpurun_out/bench_c4_v6.json; tail -3 gpurun_out/bench_c4_v6.err
python bench.py --workload C5_edits_4k --steps 20 --warmup 3 --no-extra > gpurun_out/bench_c5_v6.json 2> gpurun_out/bench_c5_v6.err; tail -c 600 gpurun_out/bench_c5_v6.json; tail -3 gpurun_out/bench_c5_v6.err
