mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --nwarm 20 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_sv.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sv_fused -s 3 -c 2 -o gpurun_out/prof_sv_r1d -f $B > gpurun_out/ncu_sv.log 2>&1
tail -5 gpurun_out/ncu_sv.log
