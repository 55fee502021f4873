mkdir -p gpurun_out
set -x
run() { # name n args
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 2952$2"
timeout 400 $TR bench.py --gpus $2 --config cfg5 --steps 3 --warmup 2 --nwarm 100 --no-cpu-baseline --no-e2e $3 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
echo "rc=$?"
}
run cfg5_8gpu 8 ""
run cfg5_8gpu_struct 8 "--structured-sv"
python - <<"PY"
import json
for f in ("cfg5_8gpu","cfg5_8gpu_struct"):
    try:
        d=json.loads([l for l in open("gpurun_out/bench_%s.json"%f) if l.startswith("{")][0]); print(f, d["n_gpus"], d["ms_per_step"], d["value"], d["phase_ms_per_step"], d["cg_iters_per_step"], d["energy_per_site"][:3], d["roofline"])
    except Exception as ex: print(f, "failed", ex); print(open("gpurun_out/bench_%s.err"%f).read()[-2500:])
PY
