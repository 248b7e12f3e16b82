"""World-size-2 `gloo` test (CPU) of the data-parallel host logic: contiguous sharding, 1/B_global
scaling, SUM all-reduce of the flat gradient (+ loss accumulators), redundant Adam -> replicas
identical and equal to the single-process step.  The arithmetic here is the oracle's (test
infrastructure); the CUDA path is checked the same way on 2 GPUs by tools/dp_check.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gmvae_b200.dist import shard_bounds
from oracle import gmvae_oracle as O


def test_shard_bounds_partition():
    for gb, w in [(10, 2), (100, 8), (7, 3), (16384 * 8, 8), (5, 8)]:
        rows = []
        for r in range(w):
            b, e = shard_bounds(gb, w, r)
            assert 0 <= b <= e <= gb
            rows += list(range(b, e))
        assert rows == list(range(gb))
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    spec = O.Spec("gmvae", data_size=48, latent_size=6, hidden_sizes=[16, 12], mixture_components=5)
    params = O.init_params(spec, seed=7)
    st = O.adam_init(params)
    GB = 11
    x, _, eps, u = O.synthetic_batch(spec, GB)
    b, e = shard_bounds(GB, world, rank)
    names = list(params)
    for step in range(2):
        terms, grads = O.loss_and_grads(spec, params, x[b:e], eps[b:e], u[b:e], global_batch=GB)
        flat = torch.cat([grads[n].reshape(-1) for n in names] + [torch.stack([terms["nll"], terms["kl_div_z"], terms["nent"]])])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)                 # what gmvae_allreduce_grads does with NCCL
        off = 0
        for n in names:
            k = grads[n].numel()
            grads[n] = flat[off:off + k].reshape(grads[n].shape); off += k
        O.adam_tf_step(params, grads, st)
    # the id exchange used for the NCCL communicator
    obj = [b"x" * 128 if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    assert obj[0] == b"x" * 128
    torch.save({"params": params, "loss": flat[off:].sum()}, out + f".{rank}")
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process(tmp_path):
    out = str(tmp_path / "dp")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    spec = O.Spec("gmvae", data_size=48, latent_size=6, hidden_sizes=[16, 12], mixture_components=5)
    params = O.init_params(spec, seed=7)
    st = O.adam_init(params)
    x, _, eps, u = O.synthetic_batch(spec, 11)
    for step in range(2):
        terms, _ = O.train_step(spec, params, st, x, eps, u)
    for n in params:
        assert torch.equal(r0["params"][n], r1["params"][n])          # replicas identical
        assert (r0["params"][n] - params[n]).abs().max() < 1e-12       # == single process
    assert abs(r0["loss"].item() - terms["loss"].item()) < 1e-10
