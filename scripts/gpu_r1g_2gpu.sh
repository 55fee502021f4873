mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -5 gpurun_out/bench_2gpu.err
python - <<"PY"
import json
d=json.loads([l for l in open("gpurun_out/bench_2gpu.json") if l.startswith("{")][0]); print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"], d["phase_ms_per_step"], d["cg_iters_per_step"], d["energy_per_site"], d["roofline"]["avg_launch_ms"])
PY
