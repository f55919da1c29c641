# time kernel variants: baseline first, then each "ENV=value" setting / VXRT_LIB=<library built elsewhere> given
set -x
python scripts/exp_time.py --workloads C3ii_4k,C3ii_pitched_4k 2>&1 | tail -3
for v in "$@"; do
  env $v python scripts/exp_time.py --workloads C3ii_4k,C3ii_pitched_4k 2>&1 | tail -3
done
