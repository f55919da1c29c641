set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -c 3500 gpurun_out/bench_c.json; tail -5 gpurun_out/bench_c.err
ncu --set full --clock-control none --import-source on -k regex:"primary_kernel|shade_kernel" -s 10 -c 2 -o gpurun_out/prof_c python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/b_ncu2.log 2>&1
tail -3 gpurun_out/b_ncu2.log
