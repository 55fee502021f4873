mkdir -p gpurun_out
for cs in 8 9 10 12 16 7 5; do
NQS_SV_CS=$cs timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3 > gpurun_out/bench_cs$cs.json 2> gpurun_out/bench_cs$cs.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_cs$cs.json") if l.startswith("{")][0]); print("cs $cs", d["roofline"]["variant"], d["roofline"]["avg_launch_ms"], d["roofline"]["achieved"], d["energy_per_site"][:2])
except Exception as ex: print("cs $cs failed", ex)
PY
done
