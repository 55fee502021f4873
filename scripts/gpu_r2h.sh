set -x
for a in "--config cfg3 --k-total 2048 --nwarm 5" "--config cfg3 --k-total 2048 --nwarm 100" "--config cfg3 --nwarm 5" "--config cfg3 --k-total 2048 --nwarm 5 --no-structured-extra"; do
timeout 400 python bench.py $a --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_dbg1.json 2> gpurun_out/bench_dbg1.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_dbg1.json") if l.startswith("{")][0]); print("$a", d["ms_per_step"], d["cg_iters_per_step"], d["energy_per_site"][:3], (d.get("structured_sv") or {}).get("energy_per_site"))
except Exception as ex: print("failed", ex); print(open("gpurun_out/bench_dbg1.err").read()[-2500:])
PY
done
