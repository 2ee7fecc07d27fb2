set -x
mkdir -p gpurun_out
for ss in 0 1; do GRB_SIDE_STREAM=$ss GRB_DEFER=1 timeout 400 python bench.py --steps 100 --skip-cpu --skip-eager --skip-roofline > gpurun_out/bench_ss$ss.json 2>gpurun_out/bench_ss$ss.err; python -c "
import json;d=json.load(open('gpurun_out/bench_ss$ss.json'));print('side_stream=$ss', d['ms_per_step'], d['value'])"; tail -2 gpurun_out/bench_ss$ss.err; done
