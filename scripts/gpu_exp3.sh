for W in 1 4 8; do for P in 0 1; do echo "== world $W prefetch $P"; python scripts/exp_time.py --world $W --prefetch $P --workloads C3ii_4k 2>&1 | grep -E "primary|rror"; done; done
