mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 240 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_2gpu_p2p.json 2> gpurun_out/bench_2gpu_p2p.err
echo "rc=$?"
python - <<"PY"
import json
for f in ("p2p",):
    try:
        d=json.loads([l for l in open("gpurun_out/bench_2gpu_%s.json"%f) if l.startswith("{")][0]); print(f, d["n_gpus"], d["ms_per_step"], d["config"]["cg_exchange"], d["phase_ms_per_step"], d["cg_iters_per_step"], d["energy_per_site"][:3], d["roofline"]["avg_launch_ms"])
    except Exception as ex: print(f, "failed", ex)
PY
