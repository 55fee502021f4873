mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --nwarm 20 --no-cpu-baseline --no-e2e --no-structured-extra"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1i.csv $B > gpurun_out/ncu_r1i_1.log 2>&1
ncu --set full --clock-control none -k regex:"sv_fused|rbm_sweep_fast|rbm_eloc_sites|cg_fused|oderiv|spin_rows_dmma|spin_cols_dmma" -s 12 -c 10 -f -o gpurun_out/prof_r1i $B > gpurun_out/ncu_r1i_3.log 2>&1
ncu -i gpurun_out/prof_r1i.ncu-rep --page raw --csv > gpurun_out/prof_r1i_raw.csv 2>/dev/null
ls -la gpurun_out/
sz=$(stat -c %s gpurun_out/prof_r1i.ncu-rep); if [ "$sz" -gt 45000000 ]; then rm gpurun_out/prof_r1i.ncu-rep; fi
tail -n 1 gpurun_out/ncu_r1i_*.log | cut -c1-160
