set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attn_tc_gpu.py -x -q --timeout 180 > gpurun_out/t_attn.txt 2>&1; tail -8 gpurun_out/t_attn.txt
timeout 400 python scripts/bench_attn.py > gpurun_out/bench_attn.jsonl 2>gpurun_out/bench_attn.err; cat gpurun_out/bench_attn.jsonl; tail -3 gpurun_out/bench_attn.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hstu_attn_tc -s 4 -c 2 -o gpurun_out/attn_tc_cfg2 -f python scripts/prof_attn.py 128 200 128 4 > gpurun_out/ncu_cfg2.log 2>&1; tail -3 gpurun_out/ncu_cfg2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hstu_attn_tc -s 4 -c 2 -o gpurun_out/attn_tc_cfg3 -f python scripts/prof_attn.py 8 2048 256 8 > gpurun_out/ncu_cfg3.log 2>&1; tail -3 gpurun_out/ncu_cfg3.log
ls -la gpurun_out/*.ncu-rep
