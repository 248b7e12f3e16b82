# Weak scaling on ONE 8-GPU box (run under gpurun --gpus 8): cfg4 per GPU, exchange over peer memory fused with Adam.
set -u
O=gpurun_out/scale2; mkdir -p $O
python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > $O/scale_n1.json 2> $O/scale_n1.err
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline > $O/scale_n$n.json 2> $O/scale_n$n.err
done
GMVAE_DP_PEER=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline > $O/scale_n8_nccl_allreduce.json 2> $O/scale_n8_nccl.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 tools/dp_check.py bf16 > $O/dp_check_n8_bf16.json 2> $O/dp_check_n8_bf16.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/dp_check.py fp32 > $O/dp_check_n8_fp32.json 2> $O/dp_check_n8_fp32.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tools/dp_check.py bf16 > $O/dp_check_n2_bf16.json 2> $O/dp_check_n2_bf16.err
# the driver's own form: 20 steps, 3 warm-up
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > $O/scale_n8_20steps.json 2> $O/scale_n8_20steps.err
ls -la $O
