set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_ddp_nccl_gpu.py -q --timeout 600 -s > gpurun_out/t_ddp.txt 2>&1; tail -40 gpurun_out/t_ddp.txt
for dp in peer nccl; do GRB_DP=$dp timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --skip-cpu --skip-eager --skip-roofline > gpurun_out/bench_n2_$dp.json 2> gpurun_out/bench_n2_$dp.err; tail -c 600 gpurun_out/bench_n2_$dp.json; tail -3 gpurun_out/bench_n2_$dp.err; done
