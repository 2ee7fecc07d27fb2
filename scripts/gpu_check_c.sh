set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attn_tc_gpu.py tests/test_hstu_gpu.py tests/test_long_seq_gpu.py -x -q --timeout 300 > gpurun_out/t_attn.txt 2>&1; tail -5 gpurun_out/t_attn.txt
timeout 400 python scripts/bench_attn.py > gpurun_out/bench_attn.jsonl 2>gpurun_out/bench_attn.err; cat gpurun_out/bench_attn.jsonl; tail -3 gpurun_out/bench_attn.err
for m in auto tc; do GRB_ATTN=$m timeout 400 python bench.py --steps 100 --skip-cpu --skip-eager > gpurun_out/bench_cfg2_$m.json 2>gpurun_out/bench_cfg2_$m.err; python -c "
import json;d=json.load(open('gpurun_out/bench_cfg2_$m.json'));print('$m', d['ms_per_step'], d['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['gpu_launches'])"; tail -2 gpurun_out/bench_cfg2_$m.err; done
for m in auto mma; do GRB_ATTN=$m timeout 900 python bench.py --config cfg3 --steps 10 --warmup 3 --skip-cpu --skip-eager > gpurun_out/bench_cfg3_$m.json 2>gpurun_out/bench_cfg3_$m.err; python -c "
import json;d=json.load(open('gpurun_out/bench_cfg3_$m.json'));print('$m', d['ms_per_step'], d['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['gpu_launches'])"; tail -2 gpurun_out/bench_cfg3_$m.err; done
