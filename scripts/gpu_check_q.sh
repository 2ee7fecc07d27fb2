mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_head_gpu.py tests/test_cfg2_parity_gpu.py tests/test_hstu_gpu.py tests/test_recall_gpu.py tests/test_sasrec_gpu.py -q --timeout 300 2>&1 | tail -12
for m in one exact; do if [ $m = one ]; then unset GRB_CE; else export GRB_CE=$m; fi; timeout 400 python bench.py --steps 100 --skip-cpu --skip-eager > gpurun_out/bench_ce_$m.json 2>gpurun_out/bench_ce_$m.err; python -c "
import json;d=json.load(open('gpurun_out/bench_ce_$m.json'));print('$m', d['ms_per_step'], d['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"; tail -2 gpurun_out/bench_ce_$m.err; done
unset GRB_CE
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_one.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-eager --no-graph > gpurun_out/b_one.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_one.csv 2>&1 | head -8
