mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_large_shapes.py -x -q 2>&1 | tail -15
for c in 0 1; do
if [ $c = 1 ]; then export NQS_SWEEP_C1=1; fi
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_c$c.json") if l.startswith("{")][0]); print("C1=$c", d["ms_per_step"], d["sweep"], d["phase_ms_per_step"], d["energy_per_site"][:2])
except Exception as ex: print("C1=$c failed", ex)
PY
done
unset NQS_SWEEP_C1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
