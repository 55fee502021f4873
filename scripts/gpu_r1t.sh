mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for sp in 1 0; do
NQS_SWEEP_SPLIT=$sp timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/bench_split$sp.json 2> gpurun_out/bench_split$sp.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_split$sp.json") if l.startswith("{")][0]); print("split $sp", d["ms_per_step"], d["sweep"], d["energy_per_site"])
except Exception as ex: print("split $sp failed", ex)
PY
done
