# Round 2, call 10 (1 GPU; call 9 stopped at a bench.py slip): the current kernels under ncu (launch list of the bench command, full capture of both render kernels on the
# benchmark frame and on C4's 1024^3 terrain), every rank's share of 2/4/8-GPU splits in isolation, the 1-GPU bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c10_pytest.log 2>&1; tail -4 gpurun_out/r2c10_pytest.log
mkdir -p gpurun_out
python scripts/rank_balance.py > gpurun_out/r2c10_rank_balance.jsonl 2>&1; cat gpurun_out/r2c10_rank_balance.jsonl | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 60 --csv --log-file gpurun_out/r2c10_launches.csv python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r2c10_ncu1.log 2>&1
tail -6 gpurun_out/r2c10_launches.csv | cut -c1-260
ncu --set full --clock-control none --import-source on -k regex:"primary_kernel|shade_kernel" -s 430 -c 2 -o gpurun_out/r2c10_prof python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r2c10_ncu2.log 2>&1
tail -2 gpurun_out/r2c10_ncu2.log
ncu --set full --clock-control none -k regex:"primary_kernel|shade_kernel" -s 430 -c 2 -o gpurun_out/r2c10_prof_C4 python bench.py --steps 10 --warmup 3 --no-extra --workload C4_terrain_4k > gpurun_out/r2c10_ncu3.log 2>&1
tail -2 gpurun_out/r2c10_ncu3.log
python bench.py --steps 40 --warmup 5 > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err; tail -c 1500 gpurun_out/r2c10_bench.json; tail -3 gpurun_out/r2c10_bench.err
python bench.py --steps 30 --warmup 5 --workload C4_terrain_4k > gpurun_out/r2c10_bench_C4.json 2> gpurun_out/r2c10_bench_C4.err; tail -c 600 gpurun_out/r2c10_bench_C4.json; tail -3 gpurun_out/r2c10_bench_C4.err
ls -la gpurun_out | tail -12
