mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:rq_residual -s 2 -c 2 -o gpurun_out/r2_rq -f python scripts/prof_rq.py > gpurun_out/prof_rq.out 2>&1; tail -1 gpurun_out/prof_rq.out
