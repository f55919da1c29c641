set -x
python scripts/exp_time.py --workloads C3ii_4k 2>&1 | tail -9
