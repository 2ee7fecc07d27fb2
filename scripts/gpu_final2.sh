set -x
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 900 ncu --metrics $M --clock-control none --cache-control none -c 700 --csv --log-file gpurun_out/r2_step_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-eager --skip-roofline > gpurun_out/prof_cfg2.out 2>&1
python scripts/step_traffic.py gpurun_out/r2_step_cfg2.csv | tail -3
timeout 600 python bench.py > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err; cut -c1-300 gpurun_out/bench_r2_default.json
timeout 900 python bench.py --config cfg3 --skip-cpu > gpurun_out/bench_r2_cfg3.json 2> gpurun_out/bench_r2_cfg3.err; cut -c1-300 gpurun_out/bench_r2_cfg3.json
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
