for P in 0 1; do echo "== world 8 prefetch $P"; python scripts/exp_time.py --world 8 --prefetch $P --workloads C3ii_4k 2>&1 | grep -E "primary|rror"; done
for P in 0 1; do echo "== world 1 prefetch $P"; python scripts/exp_time.py --world 1 --prefetch $P --workloads C3ii_4k,C2_1080p 2>&1 | grep -E "primary|rror"; done
