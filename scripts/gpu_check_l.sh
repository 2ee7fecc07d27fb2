set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fp32_gpu.py -x -q --timeout 300 2>&1 | tail -15
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -12
