# Round 2, call 5 (1 GPU): the fused frame kernel for small shares
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c5_pytest.log 2>&1; tail -4 gpurun_out/r2c5_pytest.log
python scripts/share_probe.py --world 8 > gpurun_out/r2c5_share8.jsonl 2>&1; cat gpurun_out/r2c5_share8.jsonl | cut -c1-700
python scripts/share_probe.py --world 4 > gpurun_out/r2c5_share4.jsonl 2>&1; cat gpurun_out/r2c5_share4.jsonl | cut -c1-400
python scripts/share_probe.py --world 2 > gpurun_out/r2c5_share2.jsonl 2>&1; cat gpurun_out/r2c5_share2.jsonl | cut -c1-400
for v in 0 1; do VXRT_FUSION=$v python scripts/exp_probe.py | tee -a gpurun_out/r2c5_probe_fusion.jsonl | cut -c1-600; done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or device_side or golden_frames" 2>&1 | tail -8 > gpurun_out/r2c5_memcheck.txt; tail -5 gpurun_out/r2c5_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_frames or ragged" 2>&1 | tail -6 > gpurun_out/r2c5_racecheck.txt; tail -4 gpurun_out/r2c5_racecheck.txt
