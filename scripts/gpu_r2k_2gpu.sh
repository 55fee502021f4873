mkdir -p gpurun_out
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542"
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_new.json 2> gpurun_out/bench_2gpu_new.err
timeout 300 $TR bench.py --gpus 2 --config cfg5 --k-total 4096 --structured-sv --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_2gpu_cfg5s.json 2> gpurun_out/bench_2gpu_cfg5s.err
python - <<"PY"
import json
for f in ("2gpu_new","2gpu_cfg5s"):
    try:
        d=json.loads([l for l in open("gpurun_out/bench_%s.json"%f) if l.startswith("{")][0]); print(f, d["ms_per_step"], d["instrumented_pass"]["ms_per_step"], (d["e2e"] or {}).get("ms_per_step"), d["cg_iters_per_step"], d["phase_ms_per_step"])
    except Exception as ex: print(f, "failed", ex); print(open("gpurun_out/bench_%s.err"%f).read()[-1500:])
PY
