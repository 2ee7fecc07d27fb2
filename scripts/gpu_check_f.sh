set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attn_tc_gpu.py -x -q --timeout 180 2>&1 | tail -2
timeout 400 python scripts/bench_attn.py 2>/dev/null | grep attn_tc
