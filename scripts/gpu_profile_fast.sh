mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --nwarm 20 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_fast.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"rbm_sweep_fast|rbm_eloc_fast" -s 1 -c 2 -o gpurun_out/prof_fast_r1b $B > gpurun_out/ncu_fast.log 2>&1
tail -2 gpurun_out/ncu_fast.log
