mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { # name, env, args
  env $2 timeout 600 python bench.py $3 --no-cpu-baseline --no-e2e --no-structured-extra --steps 5 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_$1.json") if l.startswith("{")][0]); print("$1", d["ms_per_step"], d["value"], d["roofline"]["avg_launch_ms"], d["phase_ms_per_step"], d["cg_iters_per_step"], d["energy_per_site"][:2])
except Exception as ex: print("$1 failed", ex); print(open("gpurun_out/bench_$1.err").read()[-1500:])
PY
}
run ov1 NQS_NO_OVERLAP=0 "--config cfg3"
run ov0 NQS_NO_OVERLAP=1 "--config cfg3"
run ov1_cfg4 NQS_NO_OVERLAP=0 "--config cfg4"
