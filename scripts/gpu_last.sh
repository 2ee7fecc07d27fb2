mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "exit $?"; wc -l gpurun_out/bench_last.json; python -c "
import json;d=json.load(open('gpurun_out/bench_last.json'));print(sorted(d.keys()));print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['cpu_baseline']['value'], d['cuda_eager_baseline']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
