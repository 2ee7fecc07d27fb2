timeout 300 python -m pytest tests/test_sasrec_gpu.py -q -k pointwise --timeout 300 2>&1 | grep -B5 -A12 "Error\|assert" | head -60
./scripts/fma_peak
