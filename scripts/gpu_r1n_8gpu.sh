mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=index,name --format=csv | head -10
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -8
for n in 8 4; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n"
timeout 300 $TR bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
echo "rc=$?"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-p2p > gpurun_out/bench_8gpu_nccl.json 2> gpurun_out/bench_8gpu_nccl.err
python - <<"PY"
import json
for f in ("8gpu","4gpu","8gpu_nccl"):
    try:
        d=json.loads([l for l in open("gpurun_out/bench_%s.json"%f) if l.startswith("{")][0]); print(f, d["n_gpus"], d["ms_per_step"], d["value"], d["config"]["cg_exchange"], (d["e2e"] or {}).get("value"), d["phase_ms_per_step"], d["cg_iters_per_step"], d["energy_per_site"][:3], d["roofline"]["avg_launch_ms"])
    except Exception as ex: print(f, "failed", ex)
PY
