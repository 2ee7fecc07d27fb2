mkdir -p gpurun_out
# one layer's worth of every block-stack kernel of a cfg-2 step, full metric set (skip the first warm-up steps)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"hstu_attn_|tc_gemm_kernel|tc_tn_group|ln_gate_|colsum|cast_" -s 200 -c 26 -o gpurun_out/r2_step_kernels_cfg2 -f python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-eager --skip-roofline > gpurun_out/prof_stepk.out 2>&1; tail -1 gpurun_out/prof_stepk.out | cut -c1-200
ls -la gpurun_out/r2_step_kernels_cfg2.ncu-rep
