# First gpurun call of the next round (one box, ~12 min): everything that was written after round 1's GPU minutes were spent,
# in the order "what the driver runs" -> "what is new".  Results under gpurun_out/.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 40 --warmup 5 > gpurun_out/bench_r2_first.json 2> gpurun_out/bench_r2_first.err; tail -c 400 gpurun_out/bench_r2_first.json; tail -3 gpurun_out/bench_r2_first.err
bash scripts/gpu_glshim.sh
bash scripts/gpu_fast_runs.sh
# per-kernel times of the experiment on both poses, twice (exp_time.py)
bash scripts/gpu_exp_variant.sh VXRT_FAST_RUNS=0 VXRT_FAST_RUNS=1
# experiment variants that change the render kernels' SASS live in libraries of their own (voxel_rt_b200.build.VARIANTS):
# the whole parity suite on the variant, then its kernel times next to the default library's
python voxel-rt_b200/build.py late_domain_check | tail -1
VXRT_LIB=$PWD/voxel-rt_b200/libvxrt_exp_late_domain_check.so python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/gpu_exp_variant.sh VXRT_FAST_RUNS=0 VXRT_LIB=voxel-rt_b200/libvxrt_exp_late_domain_check.so
python voxel-rt_b200/build.py jump_prefetch | tail -1
for v in "" "VXRT_FAST_RUNS=1" "VXRT_LIB=$PWD/voxel-rt_b200/libvxrt_exp_late_domain_check.so" "VXRT_LIB=$PWD/voxel-rt_b200/libvxrt_exp_jump_prefetch.so"; do
  env $v python scripts/exp_probe.py | tee -a gpurun_out/exp_probe.jsonl | cut -c1-400
done
