mkdir -p gpurun_out
for m in 0 1 2 6; do
NQS_SV_DEBUG=$m timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 --cg-fixed-iters 6 > gpurun_out/bench_dbg$m.json 2> gpurun_out/bench_dbg$m.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_dbg$m.json")); print("debug $m", d["roofline"]["avg_launch_ms"], d["roofline"]["achieved"])
except Exception as ex: print("debug $m failed", ex)
PY
done
