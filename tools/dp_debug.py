"""Debug helper: device-noise training under DP, loss terms of every rank every few steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import gmvae_b200
from gmvae_b200.dist import init_process_group

world, rank, local = init_process_group("nccl")
B = 16384
eng = gmvae_b200.Engine(precision="bf16", max_batch=B, device=local, seed=1234)
eng.init_data_parallel()
g = torch.Generator().manual_seed(1234 + rank)
x = (torch.rand(B, 784, generator=g) < torch.rand(784, generator=g)).to(torch.uint8).cuda()
use_graph = len(sys.argv) > 1 and sys.argv[1] == "graph"
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for _ in range(2):
        eng.train_step(x)
    if use_graph:
        eng.capture_step(x)
    for i in range(400):
        loss = eng.replay() if use_graph else eng.train_step(x)
        if i % 25 == 0 or i == 399:
            side.synchronize()
            t = loss.cpu().tolist()
            nanp = int(torch.isnan(eng.params).sum())
            nang = int(torch.isnan(eng.grads).sum())
            print(f"rank {rank} step {i} loss {t} nan_params {nanp} nan_grads {nang} tail {eng.grads[-8:].cpu().tolist()}", flush=True)
dist.barrier()
dist.destroy_process_group()
