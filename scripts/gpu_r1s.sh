mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bench_r1s.json 2> gpurun_out/bench_r1s.err; tail -3 gpurun_out/bench_r1s.err
python - <<'PY'
import json
for f in ("bench_r1s",):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["ms_per_step"], d["sweep"], d["phase_ms_per_step"], d["cg_iters_per_step"], d["energy_per_site"])
    except Exception as ex: print(f, "failed", ex)
PY
