"""Multi-GPU self-consistency; run under torchrun with N >= 2 ranks:
an N-rank step on contiguous shards must equal the 1-rank step on the concatenated batch
(loss terms, the all-reduced gradient, parameters after k Adam steps).  Rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gmvae_b200  # noqa: E402
from gmvae_b200.dist import init_process_group, shard_bounds  # noqa: E402


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    world, rank, local = init_process_group("nccl")
    GB = 4096 + 7
    cfg = dict(model="gmvae", data_size=784, latent_size=64, hidden_sizes=[512, 512], mixture_components=10)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(GB, 784, generator=g) < 0.4).to(torch.uint8)
    eps = torch.randn(GB, 64, generator=g)
    u = torch.rand(GB, 10, generator=g).clamp_min(1e-30)
    b, e = shard_bounds(GB, world, rank)
    dp = gmvae_b200.Engine(precision=precision, max_batch=e - b, device=local, seed=11, **cfg)
    dp.init_data_parallel()
    steps = 3
    for _ in range(steps):
        loss_dp = dp.train_step(x[b:e], eps=eps[b:e], gumbel_u=u[b:e], global_batch=GB)
    torch.cuda.synchronize()
    res = {"world": world, "precision": precision, "global_batch": GB}
    if rank == 0:
        one = gmvae_b200.Engine(precision=precision, max_batch=GB, device=local, seed=11, **cfg)
        for _ in range(steps):
            loss_1 = one.train_step(x, eps=eps, gumbel_u=u)
        torch.cuda.synchronize()
        l1, ld = loss_1.cpu().double(), loss_dp.cpu().double()
        res["loss_terms_rel"] = ((l1 - ld).abs() / l1.abs().clamp_min(1.0)).max().item()
        res["params_rel"] = ((one.params - dp.params).norm() / one.params.norm()).item()
        res["step"] = dp.global_step
    # replicas identical across ranks?
    chk = dp.params.double().sum().reshape(1)
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    if rank == 0:
        res["replica_checksum_spread"] = (max(v.item() for v in lst) - min(v.item() for v in lst))
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
