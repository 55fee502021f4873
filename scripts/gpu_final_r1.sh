mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err ) 2>&1 | grep real
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err ) 2>&1 | grep real
B="python bench.py --steps 2 --warmup 1 --nwarm 20 --no-cpu-baseline --no-e2e --no-structured-extra"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1i.csv $B > gpurun_out/ncu_r1i_1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1i_struct.csv $B --structured-sv > gpurun_out/ncu_r1i_2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sv_fused|rbm_sweep_fast|rbm_eloc_sites|cg_fused|oderiv|spin_rows_dmma|spin_cols_dmma" -s 10 -c 14 -f -o gpurun_out/prof_r1i $B > gpurun_out/ncu_r1i_3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"spin_rows_dmma|spin_cols_dmma" -s 8 -c 4 -f -o gpurun_out/prof_r1i_struct $B --structured-sv > gpurun_out/ncu_r1i_4.log 2>&1
tail -n 1 gpurun_out/ncu_r1i_*.log | cut -c1-200
python - <<"PY"
import json
d=json.loads([l for l in open("gpurun_out/bench_final.json") if l.startswith("{")][0])
for k in ("value","ms_per_step","e2e","gpu_launches","clocks","sweep","phase_ms_per_step","cg_iters_per_step","instrumented_pass"): print(k, d.get(k))
print({k: v for k, v in d["structured_sv"].items() if k in ("value","ms_per_step","gemm","phase_ms_per_step")})
print(d["roofline"]); print(d["cpu_baseline"])
r=json.loads([l for l in open("gpurun_out/bench_final_ref.json") if l.startswith("{")][0]); print("ref", r["value"], r["ms_per_step"], r["cpu_baseline"]["cores"], r.get("cg_iters_per_step"))
PY
