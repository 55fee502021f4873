mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --config cfg5 --k-total 8192 --structured-sv --steps 3 --warmup 2 --nwarm 100 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg5s.json 2> gpurun_out/bench_cfg5s.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_cfg5s.json") if l.startswith("{")][0]); print("cfg5 shard struct", d["ms_per_step"], d["value"], d["phase_ms_per_step"], d["cg_iters_per_step"], d["energy_per_site"][:2], d["roofline"]["avg_launch_ms"], d["roofline"]["other_pass"])
except Exception as ex: print("failed", ex); print(open("gpurun_out/bench_cfg5s.err").read()[-1500:])
PY
