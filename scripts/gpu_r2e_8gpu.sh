mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=index,name --format=csv | head -10
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4
run() { # name n args
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 2951$2"
timeout 300 $TR bench.py --gpus $2 --steps 10 --warmup 3 --no-cpu-baseline $3 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
echo "rc=$?"
}
run 8gpu 8 ""
run 8gpu_struct 8 "--structured-sv"
run 2gpu 2 "--no-e2e"
python - <<"PY"
import json
for f in ("8gpu","8gpu_struct","2gpu"):
    try:
        d=json.loads([l for l in open("gpurun_out/bench_%s.json"%f) if l.startswith("{")][0]); print(f, d["n_gpus"], d["ms_per_step"], d["value"], d["config"]["cg_exchange"], (d["e2e"] or {}).get("value"), d["phase_ms_per_step"], d["cg_iters_per_step"], d["energy_per_site"][:3], d["roofline"]["avg_launch_ms"], d["cg_ms_per_iter"])
    except Exception as ex: print(f, "failed", ex); print(open("gpurun_out/bench_%s.err"%f).read()[-1500:])
PY
