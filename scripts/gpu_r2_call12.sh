# Round 2, call 12 (1 GPU): launch-order sorts off the frame's stream (side stream, double-buffered orders), partition table; parity
# suite, the 1-GPU bench line, the slowest rank's blocks of an 8-GPU split
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c12_pytest.log 2>&1; tail -4 gpurun_out/r2c12_pytest.log
python bench.py --steps 40 --warmup 5 > gpurun_out/r2c12_bench.json 2> gpurun_out/r2c12_bench.err; tail -c 1200 gpurun_out/r2c12_bench.json; tail -3 gpurun_out/r2c12_bench.err
python scripts/share_probe.py --world 8 --rank 2 > gpurun_out/r2c12_share8_rank2.jsonl 2>&1; cat gpurun_out/r2c12_share8_rank2.jsonl | cut -c1-700
python scripts/share_probe.py --world 8 --rank 0 > gpurun_out/r2c12_share8_rank0.jsonl 2>&1; tail -2 gpurun_out/r2c12_share8_rank0.jsonl | cut -c1-700
ls -la gpurun_out | tail -6
