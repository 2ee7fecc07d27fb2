set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:tc_ce_kernel -s 2 -c 2 -o gpurun_out/r2_ce -f python bench.py --steps 1 --warmup 2 --no-graph --skip-cpu --skip-eager --skip-roofline > gpurun_out/prof_ce.out 2>&1; tail -2 gpurun_out/prof_ce.out | cut -c1-200
ls -la gpurun_out/r2_ce.ncu-rep
