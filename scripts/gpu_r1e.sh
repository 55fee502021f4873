mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_sv_fused.py -x -q 2>&1 | tail -15
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; tail -3 gpurun_out/bench_fused.err
python - <<'PY'
import json
for f in ("bench_fused",):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["ms_per_step"], d["roofline"]["variant"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["phase_ms_per_step"], d["cg_iters_per_step"])
    except Exception as ex: print(f, "failed", ex)
PY
