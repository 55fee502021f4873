mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/bench_r1u.json 2> gpurun_out/bench_r1u.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r1u.json") if l.startswith("{")][0]); print(d["ms_per_step"], d["sweep"], d["phase_ms_per_step"], d["energy_per_site"][:2])
PY
