# Round 2, call 15 (1 GPU): wide blocks for the heaviest tiles of a fused frame, step reciprocals without the range check, release
# chained behind the owner's wait; parity suite, every rank's share of an 8-GPU split with / without wide tiles, the 1-GPU line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c15_pytest.log 2>&1; tail -4 gpurun_out/r2c15_pytest.log
python scripts/rank_balance.py --worlds 8 > gpurun_out/r2c15_rank_balance_wide8.jsonl 2>&1; cat gpurun_out/r2c15_rank_balance_wide8.jsonl | cut -c1-330
VXRT_WIDE_TILES=0 python scripts/rank_balance.py --worlds 8 > gpurun_out/r2c15_rank_balance_wide0.jsonl 2>&1; cat gpurun_out/r2c15_rank_balance_wide0.jsonl | cut -c1-330
VXRT_WIDE_TILES=16 python scripts/rank_balance.py --worlds 8 > gpurun_out/r2c15_rank_balance_wide16.jsonl 2>&1; cat gpurun_out/r2c15_rank_balance_wide16.jsonl | cut -c1-330
VXRT_LIB=$PWD/voxel-rt_b200/libvxrt_exp_wide_noinline.so python scripts/rank_balance.py --worlds 8 > gpurun_out/r2c15_rank_balance_noinline.jsonl 2>&1; cat gpurun_out/r2c15_rank_balance_noinline.jsonl | cut -c1-330
python bench.py --steps 40 --warmup 5 > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err; tail -c 1500 gpurun_out/r2c15_bench.json; tail -3 gpurun_out/r2c15_bench.err
ls -la gpurun_out | tail -8
