# Round 2, call 13 (1 GPU; call 12 again: the sorts read a snapshot of the block times): launch-order sorts off the frame's stream (side stream, double-buffered orders), partition table; parity
# suite, the 1-GPU bench line, the slowest rank's blocks of an 8-GPU split
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c13_pytest.log 2>&1; tail -4 gpurun_out/r2c13_pytest.log
python bench.py --steps 40 --warmup 5 > gpurun_out/r2c13_bench.json 2> gpurun_out/r2c13_bench.err; tail -c 1200 gpurun_out/r2c13_bench.json; tail -3 gpurun_out/r2c13_bench.err
ls -la gpurun_out | tail -6
