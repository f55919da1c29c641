mkdir -p gpurun_out
ncu --set full --clock-control none --cache-control all -k regex:"primary_kernel" -s 22 -c 1 -o gpurun_out/prof_w8 python scripts/exp_time.py --world 8 --workloads C3ii_4k --frames 4 > gpurun_out/prof_w8.log 2>&1
tail -2 gpurun_out/prof_w8.log
ncu --set full --clock-control none --cache-control none -k regex:"primary_kernel" -s 22 -c 1 -o gpurun_out/prof_w8_warm python scripts/exp_time.py --world 8 --workloads C3ii_4k --frames 4 --no-flush > gpurun_out/prof_w8b.log 2>&1
tail -2 gpurun_out/prof_w8b.log
