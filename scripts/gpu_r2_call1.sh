# Round 2, first GPU call (1 GPU): what the driver runs (smoke, pytest -m gpu, bench both arms) + the ncu launch list and one
# full capture of the two frame kernels for profiles/r2_*.  Results under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi -L
free -g | head -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; tail -6 gpurun_out/r2c1_pytest.log
python bench.py --steps 40 --warmup 5 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; tail -c 1500 gpurun_out/r2c1_bench.json; tail -3 gpurun_out/r2c1_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c1_bench_ref.json 2> gpurun_out/r2c1_bench_ref.err; tail -c 600 gpurun_out/r2c1_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 60 --csv --log-file gpurun_out/r2c1_launches.csv python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r2c1_ncu1.log 2>&1
tail -8 gpurun_out/r2c1_launches.csv | cut -c1-260
ncu --set full --clock-control none --import-source on -k regex:"primary_kernel|shade_kernel" -s 430 -c 2 -o gpurun_out/r2c1_prof python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r2c1_ncu2.log 2>&1
tail -2 gpurun_out/r2c1_ncu2.log
ls -la gpurun_out | tail -12
