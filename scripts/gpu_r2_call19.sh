# Round 2, call 19 (1 GPU): the closing single-GPU run -- parity suite, smoke, the bench line (+ the reference arm), C4 and C5 lines,
# ncu launch list and full capture of the final kernels
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c19_pytest.log 2>&1; tail -4 gpurun_out/r2c19_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2c19_bench.json 2> gpurun_out/r2c19_bench.err; tail -c 900 gpurun_out/r2c19_bench.json; tail -3 gpurun_out/r2c19_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c19_bench_reference.json 2> gpurun_out/r2c19_bench_reference.err; tail -c 700 gpurun_out/r2c19_bench_reference.json; tail -2 gpurun_out/r2c19_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 60 --csv --log-file gpurun_out/r2c19_launches.csv python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r2c19_ncu1.log 2>&1
tail -4 gpurun_out/r2c19_launches.csv | cut -c1-260
ncu --set full --clock-control none --import-source on -k regex:"primary_kernel|shade_kernel" -s 430 -c 2 -o gpurun_out/r2c19_prof python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r2c19_ncu2.log 2>&1
tail -2 gpurun_out/r2c19_ncu2.log
python bench.py --steps 30 --warmup 5 --workload C4_terrain_4k > gpurun_out/r2c19_bench_C4.json 2> gpurun_out/r2c19_bench_C4.err; tail -c 400 gpurun_out/r2c19_bench_C4.json; tail -3 gpurun_out/r2c19_bench_C4.err
python bench.py --steps 30 --warmup 5 --workload C5_edits_4k > gpurun_out/r2c19_bench_C5.json 2> gpurun_out/r2c19_bench_C5.err; tail -c 400 gpurun_out/r2c19_bench_C5.json; tail -3 gpurun_out/r2c19_bench_C5.err
ls -la gpurun_out | tail -12
