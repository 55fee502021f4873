mkdir -p gpurun_out
set -x
T0=$(date +%s); python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -n 25 gpurun_out/bench_full.err | grep -E "Elapsed|Maximum resident|Error|error" 
echo "native arm wall: $(( $(date +%s) - T0 )) s"; T1=$(date +%s); python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -n 25 gpurun_out/bench_ref.err | grep -E "Elapsed|Error|error"
echo "reference arm wall: $(( $(date +%s) - T1 )) s"; cat gpurun_out/bench_ref.json | cut -c1-1500
python - <<"PY"
import json
d=json.loads([l for l in open("gpurun_out/bench_full.json") if l.startswith("{")][0])
for k in ("value","ms_per_step","e2e","roofline","cpu_baseline","clocks","gpu_launches","phase_ms_per_step","cg_iters_per_step"): print(k, d.get(k))
PY
