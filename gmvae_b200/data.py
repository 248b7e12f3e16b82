"""Input pipeline -- mirror of `create_dataset` (/root/reference/scripts/runners.py:21-62) without TFDS.

The reference loads MNIST through `tfds.load`, scales to [0,1], binarises dynamically the inverted way
(`image < uniform`, :44-47), batches, repeats and then shuffles *batches* (:50-57).  Here the raw
intensity bytes are read once from local files (IDX, optionally gzipped, or a Keras-style `mnist.npz`),
kept resident in HBM (train split: 47 MB) and every batch is binarised on the device by
`gmvae_binarize` (csrc/input.cuh): at the step's throughput a host pipeline cannot feed one GPU.
Without local files a synthetic set of the same shape stands in (BASELINE.json: no dataset download)."""
from __future__ import annotations

import gzip
import os
import struct
from typing import Callable, Iterator, Optional, Tuple

import numpy as np

IMG_SHAPE = (28, 28, 1)
SPLIT_SIZES = {"train": 60000, "test": 10000}
_IDX_FILES = {"train": ("train-images-idx3-ubyte", "train-labels-idx1-ubyte"),
              "test": ("t10k-images-idx3-ubyte", "t10k-labels-idx1-ubyte")}
_IDX_DTYPES = {0x08: np.uint8, 0x09: np.int8, 0x0B: ">i2", 0x0C: ">i4", 0x0D: ">f4", 0x0E: ">f8"}


def read_idx(path: str) -> np.ndarray:
    """One IDX file (the MNIST distribution format): magic 0x0000 <dtype> <ndim>, big-endian dims, data."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    if len(raw) < 4 or raw[0] != 0 or raw[1] != 0 or raw[2] not in _IDX_DTYPES:
        raise ValueError(f"{path}: not an IDX file")
    ndim = raw[3]
    if len(raw) < 4 + 4 * ndim:
        raise ValueError(f"{path}: truncated IDX header")
    dims = struct.unpack(">" + "I" * ndim, raw[4:4 + 4 * ndim])
    dt = np.dtype(_IDX_DTYPES[raw[2]])
    n = int(np.prod(dims)) if ndim else 1
    body = raw[4 + 4 * ndim:]
    if len(body) != n * dt.itemsize:
        raise ValueError(f"{path}: IDX payload is {len(body)} bytes, header says {n * dt.itemsize}")
    return np.frombuffer(body, dtype=dt).reshape(dims)


def write_idx(path: str, arr: np.ndarray) -> None:
    """Inverse of read_idx for uint8 arrays (used by the tests and to cache converted data)."""
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(bytes([0, 0, 0x08, arr.ndim]) + struct.pack(">" + "I" * arr.ndim, *arr.shape) + arr.tobytes())


def _find(dirname: str, stem: str) -> Optional[str]:
    for name in (stem, stem + ".gz", stem.replace("-idx", ".idx"), stem.replace("-idx", ".idx") + ".gz"):
        p = os.path.join(dirname, name)
        if os.path.exists(p):
            return p
    return None


def load_mnist(dataset_path: Optional[str], split: str) -> Optional[Tuple[np.ndarray, np.ndarray]]:
    """(intensities uint8 [N, 784], labels int64 [N]) of `split` from a directory holding the four IDX files
    (plain or .gz) or an `mnist.npz` with x_train / y_train / x_test / y_test; None when nothing is there."""
    if split not in _IDX_FILES:
        raise ValueError(f"split must be 'train' or 'test', got {split!r}")
    if not dataset_path:
        return None
    if os.path.isfile(dataset_path) and dataset_path.endswith(".npz"):
        npz = dataset_path
    else:
        npz = os.path.join(dataset_path, "mnist.npz")
        img_p, lab_p = (_find(dataset_path, s) for s in _IDX_FILES[split])
        if img_p and lab_p:
            images, labels = read_idx(img_p), read_idx(lab_p)
            if images.dtype != np.uint8 or images.ndim != 3 or labels.ndim != 1 or images.shape[0] != labels.shape[0]:
                raise ValueError(f"{img_p} / {lab_p}: unexpected shapes {images.shape} {labels.shape}")
            return images.reshape(images.shape[0], -1).copy(), labels.astype(np.int64)
    if os.path.exists(npz):
        with np.load(npz) as z:
            images, labels = z[f"x_{split}"], z[f"y_{split}"]
        if images.dtype != np.uint8:
            raise ValueError(f"{npz}: x_{split} must hold uint8 intensities")
        return images.reshape(images.shape[0], -1).copy(), labels.astype(np.int64)
    return None


def synthetic_mnist(split: str, num_examples: Optional[int] = None, data_size: int = 784) -> Tuple[np.ndarray, np.ndarray]:
    """Stand-in with MNIST's tensor contract: ten class prototypes (shared by the splits, like digits) plus
    per-example jitter, uint8 intensities [N, data_size], labels int64 [N].  Deterministic per split."""
    n = int(num_examples or SPLIT_SIZES[split])
    protos = np.random.default_rng(1234).random((10, data_size)) ** 3     # mostly dark, a few bright pixels
    rng = np.random.default_rng(2345 if split == "train" else 4321)
    labels = rng.integers(0, 10, size=n).astype(np.int64)
    jitter = rng.random((n, 1)) * 0.4 + 0.6
    images = np.clip(protos[labels] * jitter * 255.0 + 0.5, 0, 255).astype(np.uint8)
    return images, labels


class BatchSchedule:
    """The reference's visiting order (runners.py:50-57): `.batch(B)` first, so batches are contiguous runs of
    the dataset (the last one short); `.repeat()`; then `.shuffle(num_examples)` over *batches* -- with a
    buffer larger than one pass it is a random order of the batches of each pass (TF's buffer also mixes
    neighbouring passes; that detail is not reproduced).  Yields (first_row, rows)."""

    def __init__(self, num_examples: int, batch_size: int, shuffle: bool, repeat: bool, seed: Optional[int] = None):
        if num_examples <= 0 or batch_size <= 0:
            raise ValueError("num_examples and batch_size must be positive")
        self.num_examples, self.batch_size, self.shuffle, self.repeat = int(num_examples), int(batch_size), shuffle, repeat
        self.rng = np.random.default_rng(seed)
        self.num_batches = (self.num_examples + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Tuple[int, int]]:
        while True:
            order = self.rng.permutation(self.num_batches) if self.shuffle else np.arange(self.num_batches)
            for b in order.tolist():
                first = b * self.batch_size
                yield first, min(self.batch_size, self.num_examples - first)
            if not self.repeat:
                return


class DeviceDataset:
    """Intensities and labels resident on the device; iterating yields (x uint8 {0,1} [B, D], labels int64 [B]),
    both on the device, binarised freshly for every visit (dynamic binarisation, runners.py:44-47).

    `binarize(intensities, batch=, first_row=, draw=, out=)` is `Engine.binarize`; every batch gets a new draw
    counter, so no two visits of a sample share uniforms.  `static_out` (uint8 [rows, D]) makes batches of exactly
    that many rows land in one fixed buffer -- the input of a captured step graph.

    Data parallelism (`world` > 1): `batch_size` is the GLOBAL batch; every rank walks the same schedule (same
    seed) and takes its contiguous share of each global batch (dist.shard_bounds); `last_global_rows` is the
    divisor of the batch means for the step.  Global batches with fewer rows than ranks are skipped by all."""

    def __init__(self, intensities, labels, batch_size: int, shuffle: bool, repeat: bool, binarize: Callable,
                 seed: Optional[int] = None, first_draw: int = 0, static_out=None, world: int = 1, rank: int = 0):
        if intensities.shape[0] != labels.shape[0]:
            raise ValueError("intensities and labels disagree on the number of examples")
        if not (0 <= rank < world):
            raise ValueError("rank out of range")
        if world > 1 and seed is None:
            raise ValueError("data parallelism needs a schedule seed shared by all ranks")
        self.intensities, self.labels = intensities, labels
        self.schedule = BatchSchedule(intensities.shape[0], batch_size, shuffle, repeat, seed)
        self.binarize, self.draw, self.static_out = binarize, int(first_draw), static_out
        self.world, self.rank = int(world), int(rank)
        self.last_global_rows = 0

    @property
    def num_examples(self) -> int:
        return self.schedule.num_examples

    def __iter__(self):
        from .dist import shard_bounds
        for first, rows in self.schedule:
            if rows < self.world:
                continue
            b, e = shard_bounds(rows, self.world, self.rank)
            out = self.static_out if (self.static_out is not None and e - b == self.static_out.shape[0]) else None
            x = self.binarize(self.intensities, batch=e - b, first_row=first + b, draw=self.draw, out=out)
            self.draw += 1
            self.last_global_rows = rows
            yield x, self.labels[first + b:first + e]
