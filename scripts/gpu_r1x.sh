mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_structured.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sv_fused.py -x -q 2>&1 | tail -5
