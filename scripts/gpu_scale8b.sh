mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_ddp_nccl_gpu.py -q --timeout 600 > gpurun_out/t_ddp.txt 2>&1; tail -6 gpurun_out/t_ddp.txt
run() { # n dp cfg steps tag
  GRB_DP=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --config $3 --steps $4 --warmup 5 --skip-cpu --skip-eager --skip-roofline > gpurun_out/scale_$5.json 2> gpurun_out/scale_$5.err
  python -c "
import json;d=json.load(open('gpurun_out/scale_$5.json'));print('$5', d['n_gpus'], round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['value']), d['config'].get('dp_mode'))" || tail -5 gpurun_out/scale_$5.err
}
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 --skip-cpu --skip-eager --skip-roofline > gpurun_out/scale_cfg2_n1.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/scale_cfg2_n1.json'));print('cfg2_n1', round(d['ms_per_step'],4), round(d['value']))"
run 8 peer cfg2 100 cfg2_n8_peer
run 2 peer cfg2 100 cfg2_n2_peer
run 4 peer cfg2 100 cfg2_n4_peer
run 8 nccl cfg2 100 cfg2_n8_nccl
