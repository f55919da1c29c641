set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; tail -c 3500 gpurun_out/bench_d.json; tail -5 gpurun_out/bench_d.err
