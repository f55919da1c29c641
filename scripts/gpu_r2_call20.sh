# Round 2, call 20 (1 GPU): the final library (re-base order chosen per kernel) -- parity suite, bench line, ncu launch list + full
# capture of the two render kernels
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c20_pytest.log 2>&1; tail -4 gpurun_out/r2c20_pytest.log
python bench.py > gpurun_out/r2c20_bench.json 2> gpurun_out/r2c20_bench.err; tail -c 600 gpurun_out/r2c20_bench.json; tail -3 gpurun_out/r2c20_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 60 --csv --log-file gpurun_out/r2c20_launches.csv python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r2c20_ncu1.log 2>&1
tail -3 gpurun_out/r2c20_launches.csv | cut -c1-260
ncu --set full --clock-control none --import-source on -k regex:"primary_kernel|shade_kernel" -s 430 -c 2 -o gpurun_out/r2c20_prof python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r2c20_ncu2.log 2>&1
tail -2 gpurun_out/r2c20_ncu2.log
