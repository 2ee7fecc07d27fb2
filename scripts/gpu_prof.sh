set -x
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 900 ncu --metrics $M --clock-control none --cache-control none -s 380 -c 240 --csv --log-file gpurun_out/r2_step_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-eager --skip-roofline > gpurun_out/prof_cfg2.out 2>&1; tail -2 gpurun_out/prof_cfg2.out | cut -c1-200
timeout 1200 ncu --metrics $M --clock-control none --cache-control none -s 560 -c 330 --csv --log-file gpurun_out/r2_step_cfg3.csv python bench.py --config cfg3 --steps 2 --warmup 3 --no-graph --skip-cpu --skip-eager --skip-roofline > gpurun_out/prof_cfg3.out 2>&1; tail -2 gpurun_out/prof_cfg3.out | cut -c1-200
python scripts/launch_summary.py gpurun_out/r2_step_cfg2.csv | head -40
python scripts/launch_summary.py gpurun_out/r2_step_cfg3.csv | head -40
timeout 600 python -m pytest tests/test_ddp_nccl_gpu.py tests/test_cfg2_parity_gpu.py -q --timeout 600 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
