set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; free -g | head -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 3000 gpurun_out/bench_a.json; tail -5 gpurun_out/bench_a.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_a.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/b_ncu.log 2>&1
tail -20 gpurun_out/launches_a.csv
ncu --set full --clock-control none --import-source on -k regex:"primary_kernel|shade_kernel" -s 6 -c 2 -o gpurun_out/prof_a python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out
