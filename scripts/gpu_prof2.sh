set -x
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 900 ncu --metrics $M --clock-control none --cache-control none -c 700 --csv --log-file gpurun_out/r2_step_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-eager --skip-roofline > gpurun_out/prof_cfg2.out 2>&1; tail -1 gpurun_out/prof_cfg2.out | cut -c1-200
timeout 1200 ncu --metrics $M --clock-control none --cache-control none -c 1200 --csv --log-file gpurun_out/r2_step_cfg3.csv python bench.py --config cfg3 --steps 2 --warmup 3 --no-graph --skip-cpu --skip-eager --skip-roofline > gpurun_out/prof_cfg3.out 2>&1; tail -1 gpurun_out/prof_cfg3.out | cut -c1-200
python scripts/step_traffic.py gpurun_out/r2_step_cfg2.csv | tail -3
python scripts/step_traffic.py gpurun_out/r2_step_cfg3.csv | tail -3
timeout 900 ncu --set full --import-source on --clock-control none -k regex:hstu_attn_tc -s 4 -c 2 -o gpurun_out/r2_attn_tc_cfg3 -f python scripts/prof_attn.py 16 2048 256 8 > gpurun_out/prof_attn3.out 2>&1; tail -1 gpurun_out/prof_attn3.out
GRB_ATTN=mma timeout 900 ncu --set full --import-source on --clock-control none -k regex:hstu_attn_ -s 6 -c 3 -o gpurun_out/r2_attn_mma_cfg2 -f python scripts/prof_attn.py 128 200 128 4 > gpurun_out/prof_attn2.out 2>&1; tail -1 gpurun_out/prof_attn2.out
timeout 600 python bench.py > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err; cut -c1-600 gpurun_out/bench_r2_default.json
timeout 900 python bench.py --config cfg3 --skip-cpu > gpurun_out/bench_r2_cfg3.json 2> gpurun_out/bench_r2_cfg3.err; cut -c1-600 gpurun_out/bench_r2_cfg3.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; cut -c1-400 gpurun_out/bench_r2_ref.json
timeout 600 python bench.py --impl cuda_eager --steps 10 --warmup 3 > gpurun_out/bench_r2_eager.json 2> gpurun_out/bench_r2_eager.err; cut -c1-400 gpurun_out/bench_r2_eager.json
