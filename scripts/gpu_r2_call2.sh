# Round 2, second GPU call (1 GPU): the traversal grid.  Parity first (whole suite with the traversal grid, again with the plain
# kernels), then the bench line (its experiments object times with / without the traversal grid in processes of their own), the
# ncu launch list + one full capture, memcheck on the traversal tests.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c2_pytest.log 2>&1; tail -6 gpurun_out/r2c2_pytest.log
VXRT_TRAVERSAL=0 timeout 1200 python -m pytest tests -m gpu -x -q -k "not traversal and not c4_terrain and not c5_thousand" > gpurun_out/r2c2_pytest_plain.log 2>&1; tail -4 gpurun_out/r2c2_pytest_plain.log
python bench.py --steps 40 --warmup 5 > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err; tail -c 1800 gpurun_out/r2c2_bench.json; tail -3 gpurun_out/r2c2_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 60 --csv --log-file gpurun_out/r2c2_launches.csv python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r2c2_ncu1.log 2>&1
tail -6 gpurun_out/r2c2_launches.csv | cut -c1-260
ncu --set full --clock-control none --import-source on -k regex:"primary_kernel|shade_kernel" -s 430 -c 2 -o gpurun_out/r2c2_prof python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r2c2_ncu2.log 2>&1
tail -2 gpurun_out/r2c2_ncu2.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "traversal or golden_frames or ragged or known_answers or other_grid_shapes" 2>&1 | tail -12 > gpurun_out/r2c2_memcheck.txt; tail -6 gpurun_out/r2c2_memcheck.txt
for w in C3ii_pitched_4k C2_1080p; do python scripts/exp_probe.py --workload $w | tee -a gpurun_out/r2c2_probe.jsonl | cut -c1-500; VXRT_TRAVERSAL=0 python scripts/exp_probe.py --workload $w | tee -a gpurun_out/r2c2_probe.jsonl | cut -c1-500; done
ls -la gpurun_out | tail -14
