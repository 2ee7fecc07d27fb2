set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_cfg2_parity_gpu.py tests/test_hstu_gpu.py tests/test_recall_gpu.py tests/test_sasrec_gpu.py -q --timeout 300 2>&1 | tail -8
timeout 400 python bench.py --steps 100 --skip-cpu --skip-eager > gpurun_out/bench_cex.json 2>gpurun_out/bench_cex.err; python -c "
import json;d=json.load(open('gpurun_out/bench_cex.json'));print('cex', d['ms_per_step'], d['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"; tail -2 gpurun_out/bench_cex.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cex.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-eager > gpurun_out/b_cex.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_cex.csv 2>&1 | head -40
