mkdir -p gpurun_out
set -x
run() { # name n args
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 2953$2"
timeout 400 $TR bench.py --gpus $2 --config cfg5 --k-total 2048 --steps 2 --warmup 1 --nwarm 5 --no-cpu-baseline --no-e2e $3 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
echo "rc=$?"
python - <<PY
import json
f="$1"
try:
    d=json.loads([l for l in open("gpurun_out/bench_%s.json"%f) if l.startswith("{")][0]); print(f, d["n_gpus"], d["ms_per_step"], d["config"]["cg_exchange"], d["cg_iters_per_step"], d["energy_per_site"][:3])
except Exception as ex: print(f, "failed", ex); print(open("gpurun_out/bench_%s.err"%f).read()[-2500:])
PY
}
run dbg_struct 2 "--structured-sv"
run dbg_struct_nop2p 2 "--structured-sv --no-p2p"
run dbg_cfg3 2 "--config cfg3 --k-total 2048"
