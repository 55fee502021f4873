mkdir -p gpurun_out
set -x
n=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n"
NQS_CG_TRACE=1 timeout 300 $TR bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${n}gpu_trace.json 2> gpurun_out/bench_${n}gpu_trace.err
echo "rc=$?"
timeout 300 $TR bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --structured-sv > gpurun_out/bench_${n}gpu_struct.json 2> gpurun_out/bench_${n}gpu_struct.err
echo "rc=$?"
python - <<"PY"
import json
for f in ("4gpu_trace","4gpu_struct"):
    try:
        d=json.loads([l for l in open("gpurun_out/bench_%s.json"%f) if l.startswith("{")][0]); print(f, d["n_gpus"], d["ms_per_step"], d["value"], d["config"]["cg_exchange"], d["phase_ms_per_step"], d["cg_iters_per_step"], d["energy_per_site"][:3], d["roofline"]["avg_launch_ms"], d["cg_ms_per_iter"])
    except Exception as ex: print(f, "failed", ex); print(open("gpurun_out/bench_%s.err"%f).read()[-2000:])
PY
grep cgtrace gpurun_out/bench_4gpu_trace.err | grep "rank 0" | tail -16
grep cgtrace gpurun_out/bench_4gpu_trace.err | grep "rank 2" | tail -8
