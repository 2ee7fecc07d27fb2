mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_t5_attention_gpu.py -q --timeout 300 2>&1 | grep -v "^$" | tail -40
