"""Training driver -- mirror of /root/reference/scripts/runners.py `create_model` (:65-103) and
`run_train` (:106-232).  Host side only: dataset -> model -> loop; the step runs in the native
library.  Synthetic binarised 28x28 data replaces the TFDS MNIST pipeline (runners.py:21-62;
BASELINE.json: no dataset download), keeping its tensor contract: images bool [B,28,28,1],
labels int64 [B]."""
from __future__ import annotations

import json
import os
import time
from typing import Iterator, Tuple

import torch

from . import gmvae as gmvae_mod
from . import vae as vae_mod

IMG_SHAPE = (28, 28, 1)


def create_dataset(config, split: str, shuffle: bool, repeat: bool) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
    """Synthetic stand-in for runners.create_dataset: a fixed set of per-pixel intensities,
    dynamically binarised the reference's (inverted) way, `image < uniform` (runners.py:44-47)."""
    n = 60000 if split == "train" else 10000
    g = torch.Generator().manual_seed(1234 if split == "train" else 4321)
    protos = torch.rand(10, 784, generator=g)                         # one intensity pattern per class
    labels_all = torch.randint(0, 10, (n,), generator=g)
    bs = config.batch_size
    while True:
        order = torch.randperm(n // bs, generator=g) if shuffle else torch.arange(n // bs)   # batch-level shuffle (:56-57)
        for b in order.tolist():
            lab = labels_all[b * bs:(b + 1) * bs]
            inten = protos[lab]
            img = inten < torch.rand(inten.shape, generator=g)
            yield img.reshape(-1, *IMG_SHAPE), lab
        if not repeat:
            return


def create_model(config, data_dim: int):
    """runners.py:65-103: the hyper-parameters that are not flags are fixed here exactly as there."""
    hidden = [config.hidden_size] * config.num_layers
    if config.model == "gmvae":
        model = gmvae_mod.create_gmvae(data_dim, config.latent_size, mixture_components=config.mixture_components,
                                       fcnet_hidden_sizes=hidden, sigma_min=0.0, raw_sigma_bias=0.5, temperature=1.0)
    elif config.model == "vae_gmp":
        model = vae_mod.create_vae(data_dim, config.latent_size, mixture_components=config.mixture_components,
                                   fcnet_hidden_sizes=hidden, sigma_min=0.0, raw_sigma_bias=0.5)
    else:
        model = vae_mod.create_vae(data_dim, config.latent_size, fcnet_hidden_sizes=hidden, sigma_min=0.0, raw_sigma_bias=0.5)
    return model


class EarlyStopping:
    """utils.EarlyStoppingHook (utils.py:13-57): stop after `max_steps` consecutive steps without
    `loss < prev * (1 - threshold)`."""

    def __init__(self, max_steps=100, threshold=0.001):
        self.max_steps, self.threshold = max_steps, threshold
        self.steps, self.prev = 0, None

    def update(self, loss: float) -> bool:
        self.steps += 1
        if self.prev is None or loss < self.prev - self.prev * self.threshold:
            self.prev, self.steps = loss, 0
        return self.steps >= self.max_steps


def logdir_for(config) -> str:
    """runners.py:212-217."""
    return "{}/{}/h{}_n{}_z{}".format(config.logdir, config.model, config.hidden_size, config.num_layers, config.latent_size)


def run_train(config):
    """runners.py:106-232.  Differences, all forced by asynchronous execution and documented in
    DESIGN.md: the loss is read back every `summarise_every` steps (the reference's early-stopping
    hook fetches it every step, utils.py:27-30), so early stopping counts in units of that stride;
    checkpoints are `torch.save` files keyed by the reference's variable names."""
    torch.manual_seed(config.random_seed or 0)
    model = create_model(config, data_dim=784)
    model.configure(precision=getattr(config, "precision", "bf16"), objective=getattr(config, "objective", "reference"),
                    learning_rate=config.learning_rate, max_batch=config.batch_size,
                    device=int(config.gpu_id) if str(config.gpu_id).isdigit() else 0)
    model.random_seed = config.random_seed
    eng = model.engine(config.batch_size)
    logdir = logdir_for(config)
    os.makedirs(logdir, exist_ok=True)
    ckpt = os.path.join(logdir, "model.ckpt.pt")
    if os.path.exists(ckpt):                                          # MonitoredTrainingSession auto-restore
        eng.load_state_dict(torch.load(ckpt))
        print(f"Restored checkpoint at step {eng.global_step} from {ckpt}")
    data = create_dataset(config, "train", shuffle=True, repeat=True)
    stopper = EarlyStopping(max(1, config.early_stop_rounds // max(1, config.summarise_every)), config.early_stop_threshold)
    events = open(os.path.join(logdir, "summaries.jsonl"), "a")
    cur_step, last_save, t0 = eng.global_step, time.time(), time.time()
    while cur_step <= config.max_steps:
        images, labels = next(data)
        loss = eng.train_step(images.reshape(images.shape[0], -1))
        cur_step += 1
        if cur_step % config.summarise_every == 0:
            t = loss.detach().cpu().tolist()                          # the only device->host sync
            print("Step %d, %s: %f" % (cur_step, "loss", t[0]), flush=True)       # utils.py:63-65
            rec = {"step": cur_step, "elbo": -t[0], "nll_scalar": t[1], "kl_div_z": t[2],
                   "global_step/sec": config.summarise_every / max(time.time() - t0, 1e-9)}
            if config.model == "gmvae":
                rec["nent"] = t[3]
            events.write(json.dumps(rec) + "\n"); events.flush()
            t0 = time.time()
            if stopper.update(t[0]):
                print("[Early Stopping Criterion Satisfied]")
                break
        if time.time() - last_save > 120:                             # save_checkpoint_secs=120 (:226)
            torch.save(eng.state_dict(), ckpt); last_save = time.time()
    torch.save(eng.state_dict(), ckpt)
    events.close()
    return eng


def run_eval(config):
    raise NotImplementedError("--mode=eval (runners.py:235-458: t-SNE / seaborn plots) is out of scope; see DESIGN.md section 8")
