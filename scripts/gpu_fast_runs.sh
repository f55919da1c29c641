set -x
mkdir -p gpurun_out
# ray.cuh FAST_RUNS experiment: parity first, then the bench line with and without it
VXRT_TEST_EXPERIMENTS=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k fast_runs 2>&1 | tail -4
python bench.py --steps 40 --warmup 5 --no-extra > gpurun_out/bench_fast_off.json 2> gpurun_out/bench_fast_off.err
VXRT_FAST_RUNS=1 python bench.py --steps 40 --warmup 5 --no-extra > gpurun_out/bench_fast_on.json 2> gpurun_out/bench_fast_on.err
python - <<'PY'
import json
for n in ("off", "on"):
    d = json.loads(open("gpurun_out/bench_fast_%s.json" % n).read().strip().splitlines()[-1])
    print("FAST_RUNS", n, d["ms_per_step"], "ms/frame", d["roofline"]["kernels"])
PY
