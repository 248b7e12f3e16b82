"""Host-side helpers -- mirror of /root/reference/scripts/utils.py (hooks, checkpoints, image tiles,
t-SNE, clustering accuracy), without TensorFlow: arrays are numpy / torch, summaries are files.

Nothing here is on the hot path; the device work behind it is the forward-only entry points of the
native library (gmvae_encode / gmvae_decode / gmvae_prior_table)."""
from __future__ import annotations

import json
import os
import time
from typing import Optional

import numpy as np
import torch


class EarlyStoppingHook:
    """utils.py:13-57, same state machine: `after_run(loss, step)` once per training step; a stop is requested after
    `max_steps` consecutive steps without `loss < prev - prev * threshold`.  The first call (and any call after
    the global step went backwards) only resets the state, as in the reference (:38-45)."""

    def __init__(self, loss_op=None, max_steps=100, threshold=0.001):
        self._loss_op = loss_op
        self._max_steps = max_steps
        self._threshold = threshold
        self._last_step = -1
        self._steps = 0
        self._prev_loss = None
        self.stop_requested = False
        self.stop_step: Optional[int] = None

    def after_run(self, curr_loss: float, curr_step: int) -> bool:
        self._steps += 1
        if self._last_step == -1 or self._last_step > curr_step:
            self._last_step = curr_step
            self._steps = 0
            self._prev_loss = None
            return self.stop_requested
        self._last_step = curr_step
        if self._prev_loss is None or curr_loss < (self._prev_loss - self._prev_loss * self._threshold):
            self._prev_loss = curr_loss
            self._steps = 0
        if self._steps >= self._max_steps and not self.stop_requested:
            print("[Early Stopping Criterion Satisfied]", flush=True)
            self.stop_requested, self.stop_step = True, curr_step
        return self.stop_requested


def summary_formatter(log_dict) -> str:
    """utils.py:63-65: the line the reference's LoggingTensorHook prints."""
    return "Step %d, %s: %f" % (log_dict["step"], "loss", log_dict["loss"])


# ---------------------------------------------------------------------------- checkpoints
# MonitoredTrainingSession / Saver layout (runners.py:222-228): `model.ckpt-<step>` files plus a `checkpoint`
# state file naming the latest one; at most `max_to_keep` (Saver default 5) are kept.  The payload is
# Engine.state_dict() (the reference's variable names, `<var>/Adam`, `<var>/Adam_1`, `global_step`) saved by torch.
CKPT_PREFIX = "model.ckpt"


def save_checkpoint(state_dict, logdir: str, step: int, max_to_keep: int = 5) -> str:
    os.makedirs(logdir, exist_ok=True)
    name = f"{CKPT_PREFIX}-{int(step)}"
    tmp = os.path.join(logdir, name + ".tmp")
    torch.save(state_dict, tmp)
    os.replace(tmp, os.path.join(logdir, name))
    state = get_checkpoint_state(logdir) or {"model_checkpoint_path": None, "all_model_checkpoint_paths": []}
    paths = [p for p in state["all_model_checkpoint_paths"] if p != name] + [name]
    for old in paths[:-max_to_keep] if max_to_keep > 0 else []:
        try:
            os.remove(os.path.join(logdir, old))
        except OSError:
            pass
    paths = paths[-max_to_keep:] if max_to_keep > 0 else paths
    with open(os.path.join(logdir, "checkpoint.tmp"), "w") as f:
        f.write(f'model_checkpoint_path: "{name}"\n')
        for p in paths:
            f.write(f'all_model_checkpoint_paths: "{p}"\n')
    os.replace(os.path.join(logdir, "checkpoint.tmp"), os.path.join(logdir, "checkpoint"))
    return os.path.join(logdir, name)


def get_checkpoint_state(logdir: str):
    """tf.train.get_checkpoint_state: parses `<logdir>/checkpoint`; None when there is none."""
    path = os.path.join(logdir, "checkpoint")
    if not os.path.exists(path):
        return None
    latest, all_paths = None, []
    for line in open(path):
        key, _, val = line.partition(":")
        val = val.strip().strip('"')
        if key.strip() == "model_checkpoint_path":
            latest = val
        elif key.strip() == "all_model_checkpoint_paths":
            all_paths.append(val)
    if latest is None:
        return None
    return {"model_checkpoint_path": latest, "all_model_checkpoint_paths": all_paths}


def restore_checkpoint_if_exists(engine, logdir: str) -> bool:
    """utils.py:76-98 with the engine in the role of (saver, sess)."""
    state = get_checkpoint_state(logdir)
    if not state:
        return False
    full = os.path.join(logdir, os.path.basename(state["model_checkpoint_path"]))
    engine.load_state_dict(torch.load(full))
    return True


def wait_for_checkpoint(engine, logdir: str, poll_secs: float = 60.0, max_wait: Optional[float] = None) -> None:
    """utils.py:101-111: loops until a checkpoint could be restored (`max_wait` is additive: None = forever)."""
    t0 = time.time()
    while not restore_checkpoint_if_exists(engine, logdir):
        if max_wait is not None and time.time() - t0 >= max_wait:
            raise TimeoutError(f"no checkpoint appeared in {logdir} within {max_wait} s")
        print("Checkpoint not found in %s, sleeping for %g seconds." % (logdir, poll_secs), flush=True)
        time.sleep(poll_secs)


# ---------------------------------------------------------------------------- summaries
class SummaryWriter:
    """Stand-in for the TF event file: one JSON object per line (`summaries.jsonl`), images as PNG files."""

    def __init__(self, logdir: str):
        os.makedirs(logdir, exist_ok=True)
        self.logdir = logdir
        self._f = open(os.path.join(logdir, "summaries.jsonl"), "a")

    def scalars(self, step: int, values: dict) -> None:
        self._f.write(json.dumps({"step": int(step), **{k: float(v) for k, v in values.items()}}) + "\n")
        self._f.flush()

    def image(self, name: str, step: int, tile: np.ndarray) -> str:
        d = os.path.join(self.logdir, "image_summaries", name)
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, f"step_{int(step)}.png")
        save_png(path, tile)
        return path

    def close(self) -> None:
        self._f.close()


def save_png(path: str, img: np.ndarray) -> None:
    """img: [H, W] or [H, W, 1|3], floats in [0,1] or uint8."""
    from PIL import Image
    a = np.asarray(img)
    if a.ndim == 4:
        a = a[0]
    if a.ndim == 3 and a.shape[-1] == 1:
        a = a[..., 0]
    if a.dtype != np.uint8:
        a = (np.clip(a.astype(np.float64), 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
    Image.fromarray(a).save(path if path.endswith(".png") else path + ".png")


def pack_images(images, rows: int, cols: int) -> np.ndarray:
    """utils.py:114-133: a [1, rows*W, cols*H, depth] field of the first rows*cols images."""
    images = np.asarray(images)
    width, height, depth = images.shape[-3], images.shape[-2], images.shape[-1]
    images = images.reshape(-1, width, height, depth)
    batch = images.shape[0]
    rows = min(rows, batch)
    cols = min(batch // rows, cols)
    images = images[:rows * cols].reshape(rows, cols, width, height, depth)
    images = images.transpose(0, 2, 1, 3, 4)
    return images.reshape(1, rows * width, cols * height, depth)


def image_tile_summary(writer: SummaryWriter, name: str, tensor, step: int, rows: int = 8, cols: int = 8) -> str:
    """utils.py:136-137 (tf.summary.image of the packed tile, max_outputs=1)."""
    return writer.image(name, step, pack_images(_to_numpy(tensor).astype(np.float32), rows, cols))


def _to_numpy(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def flatten_tensor(inputs, shape, name="flattened"):
    """utils.py:140-141."""
    return inputs.reshape(-1, int(np.prod(shape)))


def unflatten_tensor(inputs, shape, name="unflattened"):
    """utils.py:144-145."""
    return inputs.reshape(-1, shape[0], shape[1], shape[2])


def reduce_dimensionality(data, dim: int = 2, perplexity: float = 40, random_state=None):
    """utils.py:148-153: t-SNE (300 iterations) when the data has more than two columns."""
    data = _to_numpy(data)
    if data.shape[-1] > 2:
        import inspect
        from sklearn.manifold import TSNE
        iters = "max_iter" if "max_iter" in inspect.signature(TSNE.__init__).parameters else "n_iter"
        tsne = TSNE(n_components=dim, verbose=0, perplexity=min(perplexity, max(1.0, (data.shape[0] - 1) / 3.0)),
                    random_state=random_state, **{iters: 300})
        data = tsne.fit_transform(data)
    return data


def mode_tensor(x) -> float:
    """utils.py:156-162: tf.unique_with_counts keeps first-occurrence order and argmax takes the first maximum,
    so ties go to the value that appears first."""
    x = _to_numpy(x).reshape(-1)
    vals, first, counts = np.unique(x, return_index=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    vals, counts = vals[order], counts[order]
    return float(vals[int(np.argmax(counts))])


def entropy(logits, targets):
    """utils.py:165-170: -sum(targets * log_softmax(logits), axis=1)."""
    logits = torch.as_tensor(logits)
    return -(torch.as_tensor(targets) * torch.log_softmax(logits, dim=1)).sum(dim=1)


def cluster_acc(logits, labels, no_components: int) -> float:
    """utils.py:173-191 (with `range` for the reference's Python-2 `xrange`): every predicted cluster is mapped to
    the most frequent true label among its members; accuracy of that mapping."""
    logits, labels = _to_numpy(logits), _to_numpy(labels).reshape(-1)
    cat_preds = np.argmax(logits, axis=1)
    real_preds = np.zeros(cat_preds.shape, dtype=np.float32)
    for k in range(no_components):
        idx = cat_preds == k
        lab = labels[idx]
        mode = 0.0 if lab.size == 0 else mode_tensor(lab)
        real_preds += idx.astype(np.float32) * np.float32(mode)
    return float(np.mean((real_preds == labels.astype(np.float32)).astype(np.float32)))


# ---------------------------------------------------------------------------- plots of --mode=eval (matplotlib / seaborn are absent)
_HLS = None


def _palette(n: int) -> np.ndarray:
    """n evenly spaced hues (seaborn's 'hls' palette: l = 0.6, s = 0.65)."""
    import colorsys
    return np.array([[int(255 * c) for c in colorsys.hls_to_rgb(i / max(n, 1), 0.6, 0.65)] for i in range(n)], dtype=np.uint8)


def scatter_png(path: str, xy, labels=None, n_classes: int = 10, size: int = 800, radius: int = 2) -> None:
    """A plain raster scatter plot (runners.plot_latent / plot_prior_samples, runners.py:362-394)."""
    xy = _to_numpy(xy).astype(np.float64).reshape(-1, 2)
    img = np.full((size, size, 3), 255, dtype=np.uint8)
    if xy.shape[0]:
        lo, hi = xy.min(axis=0), xy.max(axis=0)
        span = np.where(hi - lo > 0, hi - lo, 1.0)
        pix = ((xy - lo) / span * (size - 1 - 2 * radius) + radius).astype(np.int64)
        cols = _palette(n_classes)[np.asarray(labels).reshape(-1).astype(np.int64) % n_classes] if labels is not None \
            else np.tile(np.array([[40, 80, 160]], dtype=np.uint8), (xy.shape[0], 1))
        for dx in range(-radius, radius + 1):
            for dy in range(-radius, radius + 1):
                px = np.clip(pix[:, 0] + dx, 0, size - 1)
                py = np.clip(size - 1 - pix[:, 1] + dy, 0, size - 1)
                img[py, px] = cols
    save_png(path, img)


def display_images(path: str, images, n: int = 10) -> int:
    """runners.display_images (runners.py:397-410): an n x n grid of the first n*n images; with fewer images the
    grid shrinks (the reference would raise IndexError).  Returns the grid side used."""
    images = _to_numpy(images)
    images = images.reshape(images.shape[0], images.shape[1], images.shape[2], -1)
    n = max(1, min(n, int(np.floor(np.sqrt(images.shape[0])))))
    save_png(path, pack_images(images[:n * n], n, n))
    return n
