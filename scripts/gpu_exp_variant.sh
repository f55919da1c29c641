# time kernel variants: each argument is "ENV=value" (e.g. VXRT_LIB=voxel-rt_b200/libexp_x.so, relative to the repo root)
set -x
for rep in 1 2; do
for v in "$@"; do
  case "$v" in VXRT_LIB=*) v="VXRT_LIB=$PWD/${v#VXRT_LIB=}";; esac
  env $v python scripts/exp_time.py --workloads C3ii_4k,C3ii_pitched_4k 2>&1 | tail -2
done
done
