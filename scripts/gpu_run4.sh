set -x
mkdir -p gpurun_out
python bench.py --workload C4_terrain_4k --steps 20 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 2500 gpurun_out/bench_c4.json; tail -5 gpurun_out/bench_c4.err
python bench.py --workload C5_edits_4k --steps 30 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 2500 gpurun_out/bench_c5.json; tail -5 gpurun_out/bench_c5.err
