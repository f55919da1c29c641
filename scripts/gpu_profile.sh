set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"primary_kernel|shade_kernel" -s 430 -c 2 -o gpurun_out/prof_e python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/b_ncu3.log 2>&1
tail -2 gpurun_out/b_ncu3.log
