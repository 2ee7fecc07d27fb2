mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_hstu_gpu.py tests/test_cfg2_parity_gpu.py tests/test_sasrec_gpu.py tests/test_fp32_gpu.py tests/test_rq_gpu.py -q --timeout 300 -x 2>&1 | tail -5
timeout 400 python bench.py --steps 100 --skip-cpu --skip-eager > gpurun_out/bench_gemm.json 2>gpurun_out/bench_gemm.err; python -c "
import json;d=json.load(open('gpurun_out/bench_gemm.json'));print('bench', d['ms_per_step'], d['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"; tail -2 gpurun_out/bench_gemm.err
