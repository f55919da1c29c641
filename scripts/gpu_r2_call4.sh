# Round 2, call 4 (1 GPU): new defaults (primary pass on the plain grid for whole frames, divide-first re-base), diagnostics of one
# rank's share of an 8- and 4-GPU split
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c4_pytest.log 2>&1; tail -4 gpurun_out/r2c4_pytest.log
python scripts/share_probe.py --world 8 > gpurun_out/r2c4_share8.jsonl 2>&1; cat gpurun_out/r2c4_share8.jsonl | cut -c1-900
python scripts/share_probe.py --world 4 > gpurun_out/r2c4_share4.jsonl 2>&1; cat gpurun_out/r2c4_share4.jsonl | cut -c1-900
python bench.py --steps 40 --warmup 5 > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err; tail -c 2500 gpurun_out/r2c4_bench.json; tail -3 gpurun_out/r2c4_bench.err
