set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/t_all.txt 2>&1; tail -30 gpurun_out/t_all.txt
cat gpurun_out/error_table_cfg2.md
