mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_sv_fused.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
for c in cfg2 cfg3; do
for d in 1 2 3; do
NQS_SV_DEPTH=$d timeout 300 python bench.py --config $c --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/bench_${c}_d$d.json 2> gpurun_out/bench_${c}_d$d.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${c}_d$d.json") if l.startswith("{")][0]); print("$c depth $d", d["ms_per_step"], d["roofline"]["variant"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["energy_per_site"][:2])
except Exception as ex: print("$c failed", ex)
PY
done
done
