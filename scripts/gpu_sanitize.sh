set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_frames or ragged or known_answers or destroy_sequence or update_partial or peer_memory or banded or pipelined or miss_culling or terrain" 2>&1 | tail -15 > gpurun_out/sanitize_memcheck.txt; tail -8 gpurun_out/sanitize_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_frames or ragged" 2>&1 | tail -8 > gpurun_out/sanitize_racecheck.txt; tail -5 gpurun_out/sanitize_racecheck.txt
