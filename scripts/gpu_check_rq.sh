mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rq_gpu.py tests/test_sasrec_gpu.py -q --timeout 300 2>&1 | tail -4
timeout 600 python scripts/bench_kernels.py 2>/dev/null | grep rq_residual | grep '"auto"\|"tile"' | cut -c1-200 | tee gpurun_out/r2_microbench_rq2.jsonl
