mkdir -p gpurun_out
set -x
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
for c in cfg1 cfg2 cfg4; do
timeout 600 python bench.py --config $c --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_$c.json") if l.startswith("{")][0]); print("$c", d["ms_per_step"], d["value"], d["sweep"]["kernel"], d["roofline"]["variant"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["phase_ms_per_step"], d["cg_iters_per_step"], d["energy_per_site"][:2])
except Exception as ex: print("$c failed", ex); print(open("gpurun_out/bench_$c.err").read()[-1500:])
PY
done
