set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 40 --warmup 5 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; tail -c 600 gpurun_out/bench_f.json; tail -3 gpurun_out/bench_f.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 900 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 60 --csv --log-file gpurun_out/launches_f.csv python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/b_ncu_f.log 2>&1
tail -12 gpurun_out/launches_f.csv | cut -c1-260
ncu --set full --clock-control none --import-source on -k regex:"primary_kernel|shade_kernel" -s 430 -c 2 -o gpurun_out/prof_f python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/b_ncu_f2.log 2>&1
tail -2 gpurun_out/b_ncu_f2.log
