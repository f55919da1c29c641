set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 40 --warmup 5 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; tail -c 400 gpurun_out/bench_g.json; tail -3 gpurun_out/bench_g.err
bash scripts/gpu_sanitize.sh
