mkdir -p gpurun_out
timeout 600 python scripts/bench_kernels.py > gpurun_out/r2_microbench.jsonl 2> gpurun_out/r2_microbench.err; cut -c1-330 gpurun_out/r2_microbench.jsonl
timeout 900 python -m pytest tests/test_fp32_gpu.py tests/test_cfg2_parity_gpu.py tests/test_rq_gpu.py -q --timeout 300 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
