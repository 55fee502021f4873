mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --nwarm 20 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_r1e3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"rbm_sweep_fast|rbm_eloc_sites|setup_structured|oderiv" -s 1 -c 4 -f -o gpurun_out/prof_r1e_sampler $B > gpurun_out/ncu_r1e_3.log 2>&1
tail -n 3 gpurun_out/ncu_r1e_3.log | cut -c1-300
python - <<"PY"
import json
d=json.loads([l for l in open("gpurun_out/plain_r1e3.log") if l.startswith("{")][0]); print(d["ms_per_step"], d["phase_ms_per_step"], d["cg_iters_per_step"])
PY
