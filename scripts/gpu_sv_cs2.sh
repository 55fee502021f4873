mkdir -p gpurun_out
for cs in 8 9; do
NQS_SV_CS=$cs timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 2 --cg-fixed-iters 50 > gpurun_out/bench_cs$cs.json 2> gpurun_out/bench_cs$cs.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_cs$cs.json") if l.startswith("{")][0]); p=d["phase_ms_per_step"]; print("cs $cs", d["roofline"]["variant"], "sv", d["roofline"]["avg_launch_ms"], "cg_ms", p["cg_ms"], "rows", p["rows_ms"], "per-iter total", p["cg_ms"]/51, "overhead/iter", (p["cg_ms"]-p["rows_ms"])/51)
except Exception as ex: print("cs $cs failed", ex)
PY
done
