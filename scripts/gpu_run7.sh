set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; tail -c 3800 gpurun_out/bench_e.json; tail -5 gpurun_out/bench_e.err
