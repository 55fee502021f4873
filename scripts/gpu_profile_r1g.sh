mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --nwarm 20 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_r1g.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1g.csv $B > gpurun_out/ncu_r1g_1.log 2>&1
$B > gpurun_out/plain_r1g2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"sv_fused|spin_rows_dmma|cg_fused|rbm_eloc_sites" -s 8 -c 8 -f -o gpurun_out/prof_r1g $B > gpurun_out/ncu_r1g_2.log 2>&1
$B --structured-sv > gpurun_out/plain_r1g3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1g_struct.csv $B --structured-sv > gpurun_out/ncu_r1g_3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"spin_rows_dmma|spin_cols_dmma|hidden_values" -s 6 -c 6 -f -o gpurun_out/prof_r1g_struct $B --structured-sv > gpurun_out/ncu_r1g_4.log 2>&1
tail -n 2 gpurun_out/ncu_r1g_1.log gpurun_out/ncu_r1g_2.log gpurun_out/ncu_r1g_3.log gpurun_out/ncu_r1g_4.log | cut -c1-300
ls -la gpurun_out/*r1g*
