mkdir -p gpurun_out
set -x
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -4
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | tail -4
python - <<"PY"
import json
d=json.loads([l for l in open("gpurun_out/bench_default.json") if l.startswith("{")][0])
for k in ("value","ms_per_step","e2e","gpu_launches","clocks","sweep","phase_ms_per_step","cg_iters_per_step","structured_sv","cpu_baseline"): print(k, d.get(k))
print(d["roofline"])
r=json.loads([l for l in open("gpurun_out/bench_ref.json") if l.startswith("{")][0]); print(r["value"], r["ms_per_step"], r["cpu_baseline"], r.get("cg_iters_per_step"))
PY
tail -5 gpurun_out/bench_default.err
