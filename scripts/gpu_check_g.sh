set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cfg2_parity_gpu.py tests/test_attn_tc_gpu.py tests/test_hstu_gpu.py tests/test_recall_gpu.py -q --timeout 300 2>&1 | tail -4
for d in 0 1; do GRB_DEFER=$d timeout 400 python bench.py --steps 100 --skip-cpu --skip-eager > gpurun_out/bench_defer$d.json 2>gpurun_out/bench_defer$d.err; python -c "
import json;d=json.load(open('gpurun_out/bench_defer$d.json'));print('defer=$d', d['ms_per_step'], d['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"; tail -2 gpurun_out/bench_defer$d.err; done
GRB_DEFER=1 timeout 900 python bench.py --config cfg3 --steps 10 --warmup 3 --skip-cpu --skip-eager > gpurun_out/bench_cfg3_defer1.json 2>gpurun_out/bench_cfg3_defer1.err; python -c "
import json;d=json.load(open('gpurun_out/bench_cfg3_defer1.json'));print('cfg3 defer=1', d['ms_per_step'], d['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"; tail -2 gpurun_out/bench_cfg3_defer1.err
