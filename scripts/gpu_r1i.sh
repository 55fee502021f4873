mkdir -p gpurun_out
set -x
NQS_SV_DEFER=1 timeout 600 python -m pytest tests/test_gpu_sv_fused.py -x -q 2>&1 | tail -8
for d in 0 1; do
NQS_SV_DEFER=$d timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/bench_defer$d.json 2> gpurun_out/bench_defer$d.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_defer$d.json") if l.startswith("{")][0]); print("defer $d", d["ms_per_step"], d["roofline"]["variant"], d["roofline"]["avg_launch_ms"], d["roofline"]["achieved"])
except Exception as ex: print("defer $d failed", ex)
PY
done
timeout 900 python -m pytest tests/test_gpu_convergence.py -x -q 2>&1 | tail -8
