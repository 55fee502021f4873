mkdir -p gpurun_out
set -x
run() { # name, env, args
  env $2 timeout 600 python bench.py $3 --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_$1.json") if l.startswith("{")][0]); print("$1", d["ms_per_step"], d["value"], d["sweep"]["ms"], d["phase_ms_per_step"]["eloc_ms"], d["energy_per_site"][:2])
except Exception as ex: print("$1 failed", ex); print(open("gpurun_out/bench_$1.err").read()[-1500:])
PY
}
run e4 NQS_ELOC_C=4 "--config cfg3 --structured-sv"
run e8 NQS_ELOC_C=8 "--config cfg3 --structured-sv"
run e2 NQS_ELOC_C=2 "--config cfg3 --structured-sv"
