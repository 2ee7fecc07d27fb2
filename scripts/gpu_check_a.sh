set -x
mkdir -p gpurun_out
timeout 60 ./scripts/umma_probe > gpurun_out/probe.txt 2>&1; tail -12 gpurun_out/probe.txt
timeout 900 python -m pytest tests/test_attn_tc_gpu.py -x -q --timeout 180 > gpurun_out/t_attn.txt 2>&1; tail -25 gpurun_out/t_attn.txt
timeout 900 python -m pytest tests/test_cfg2_parity_gpu.py tests/test_rq_gpu.py tests/test_recall_gpu.py -q --timeout 300 > gpurun_out/t_parity.txt 2>&1; tail -25 gpurun_out/t_parity.txt
timeout 400 python scripts/bench_attn.py > gpurun_out/bench_attn.jsonl 2>gpurun_out/bench_attn.err; cat gpurun_out/bench_attn.jsonl; tail -3 gpurun_out/bench_attn.err
timeout 400 python bench.py --steps 50 --skip-cpu --skip-eager > gpurun_out/bench_quick.json 2>gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
