set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/t_all.txt 2>&1; tail -4 gpurun_out/t_all.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json | cut -c1-400
timeout 300 python bench.py --impl cuda_eager --steps 5 --warmup 2 > gpurun_out/bench_eager.json 2>gpurun_out/bench_eager.err; cat gpurun_out/bench_eager.json | cut -c1-500; tail -2 gpurun_out/bench_eager.err
