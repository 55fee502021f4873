mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --nwarm 20 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_r1e.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1e.csv $B > gpurun_out/ncu_r1e_1.log 2>&1
$B > gpurun_out/plain_r1e2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"sv_fused|rbm_sweep_fast|rbm_eloc_sites|setup_structured|cg_fused|oderiv" -s 6 -c 8 -f -o gpurun_out/prof_r1e $B > gpurun_out/ncu_r1e_2.log 2>&1
tail -3 gpurun_out/ncu_r1e_1.log gpurun_out/ncu_r1e_2.log | cut -c1-300
