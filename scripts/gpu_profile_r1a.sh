mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
B="python bench.py --steps 2 --warmup 1 --nwarm 5 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1a.csv $B > gpurun_out/ncu1.log 2>&1
$B > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"matvec_(rows|cols)" -s 4 -c 4 -o gpurun_out/prof_matvec_r1a $B > gpurun_out/ncu2.log 2>&1
$B > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"sweep_generic|eloc_generic|oderiv|setup_partial" -s 2 -c 5 -o gpurun_out/prof_other_r1a $B > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
ls -la gpurun_out
