for T in 256 128 64; do echo "== shade threads $T"; VXRT_SHADE_THREADS=$T python scripts/exp_time.py 2>&1 | grep -E "primary|rror"; done
