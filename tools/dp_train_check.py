"""run_gmvae --mode=train under torchrun (one process per GPU): every rank walks the same batch schedule, takes its
shard of each global batch, the native step all-reduces the gradients, and the replicas must stay bit-identical.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_train_check.py

Prints one JSON line on rank 0."""
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gmvae_b200 import data, run_gmvae, runners  # noqa: E402


def main():
    data.SPLIT_SIZES = {"train": 8192, "test": 512}                   # synthetic stand-in, small
    rank = int(os.environ.get("RANK", "0"))
    logdir = os.environ.get("DP_CHECK_LOGDIR") or os.path.join(tempfile.gettempdir(), "gmvae_dp_check")
    cfg = run_gmvae.build_parser().parse_args(
        ["--mode=train", "--model=gmvae", "--latent_size=16", "--hidden_size=128", "--num_layers=2", "--batch_size=256",
         f"--logdir={logdir}", "--random_seed=11", "--summarise_every=20", "--max_steps=59"])
    eng = runners.run_train(cfg)
    world = dist.get_world_size() if dist.is_initialized() else 1
    p = eng.params.double()
    stats = torch.stack([p.sum(), (p * p).sum(), torch.tensor(float(eng.global_step), device=p.device, dtype=torch.float64)])
    lo, hi = stats.clone(), stats.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out = {"world": world, "global_step": eng.global_step, "replica_checksum_spread": float((hi - lo).abs().max()),
           "param_sum": float(stats[0]), "global_batch": 256 * world}
    if rank == 0:
        recs = [json.loads(l) for l in open(os.path.join(runners.logdir_for(cfg), "summaries.jsonl"))]
        out["elbo_by_summary"] = [round(r["elbo"], 3) for r in recs]
        out["steps_summarised"] = [r["step"] for r in recs]
        print(json.dumps(out), flush=True)
    ok = out["replica_checksum_spread"] == 0.0 and eng.global_step == 60
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
