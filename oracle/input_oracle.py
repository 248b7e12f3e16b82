"""CPU oracle of the input pipeline's per-sample arithmetic -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product
(gmvae_b200/) never does.

What it restates (numpy, bit-exact integer / IEEE-fp32 arithmetic):

* `binarize`  -- runners.create_dataset._preprocess, /root/reference/scripts/runners.py:44-47:
      image = tf.cast(sample['image'], tf.float32) / 255.
      image = image < tf.random.uniform(tf.shape(image))
  The reference draws the uniforms from TF's stateful generator (not reproducible); here
  they come from a counter-based generator so that the CUDA kernel and this oracle see the same
  draws.  Parity of the *comparison* is bit-exact; parity of the *distribution* with the
  reference's is the statement P(x = 1) = 1 - intensity/255 (checked statistically in the tests).
* `philox4x32_10` -- Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3"
  (SC'11), the generator of the Random123 library; not part of the reference (third-party
  algorithm, restated from the paper).  Pinned by the Random123 known-answer vectors
  (tests/test_input_cpu.py).
* `batch_order` / `epoch_batches` -- the reference's "batch, then shuffle" order
  (runners.py:50-57: `.batch(batch_size)` precedes `.shuffle(num_examples)`, so whole batches
  are shuffled, samples inside a batch stay in dataset order, and the final short batch exists).

Parity unpinned against the reference itself (TensorFlow / TFDS cannot be installed here).
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)
DRAW_MIX = 0x9E3779B97F4A7C15
BINARIZE_STREAM = 0x8000000000000000


def philox4x32_10(counter, key):
    """counter: uint32 array [..., 4]; key: (k0, k1) python ints.  Returns uint32 [..., 4]."""
    c = [np.asarray(counter)[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[0]                      # 32x32 -> 64 bit products (no overflow in uint64)
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def philox_words(seed: int, stream: int, counters) -> np.ndarray:
    """The library's keying: counter = (ctr_lo, ctr_hi, stream_lo, stream_hi), key = (seed_lo, seed_hi)."""
    ctr = np.asarray(counters, dtype=np.uint64)
    stream = int(stream) & (2 ** 64 - 1)
    c = np.stack([(ctr & MASK32), (ctr >> np.uint64(32)),
                  np.full_like(ctr, stream & 0xFFFFFFFF), np.full_like(ctr, stream >> 32)], axis=-1).astype(np.uint32)
    seed = int(seed) & (2 ** 64 - 1)
    return philox4x32_10(c, (seed & 0xFFFFFFFF, seed >> 32))


def u01(words) -> np.ndarray:
    """24-bit uniform in the open interval (0,1): ((r >> 9) + 0.5) * 2^-23, exact in fp32."""
    return ((np.asarray(words, dtype=np.uint32) >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)


def uniforms(seed: int, draw: int, rank: int, n: int) -> np.ndarray:
    """The n uniforms the binarisation of n output bytes consumes (element e uses word e % 4 of call e // 4)."""
    key = (int(seed) ^ ((int(draw) * DRAW_MIX) & (2 ** 64 - 1))) & (2 ** 64 - 1)
    q = np.arange((n + 3) // 4, dtype=np.uint64)
    w = philox_words(key, BINARIZE_STREAM + int(rank), q).reshape(-1)
    return u01(w[:n])


def binarize(intensities: np.ndarray, row_index, batch: int, seed: int, draw: int, rank: int = 0) -> np.ndarray:
    """runners.py:44-47 on bytes: x[r, d] = (float32(intensity[src(r), d]) / 255 < u[r, d]).  uint8 {0,1} [batch, D]."""
    inten = np.asarray(intensities, dtype=np.uint8)
    D = inten.shape[1]
    rows = np.arange(batch) if row_index is None else np.asarray(row_index, dtype=np.int64)
    src = inten[rows]
    unit = src.astype(np.float32) / np.float32(255.0)                     # IEEE fp32 division, like tf.cast(...)/255.
    u = uniforms(seed, draw, rank, batch * D).reshape(batch, D)
    return (unit < u).astype(np.uint8)


def batch_order(num_examples: int, batch_size: int, rng: np.random.Generator, shuffle: bool) -> np.ndarray:
    """Order in which the batches of one pass are visited (runners.py:50-57)."""
    nb = (num_examples + batch_size - 1) // batch_size
    return rng.permutation(nb) if shuffle else np.arange(nb)


def epoch_batches(num_examples: int, batch_size: int, order) -> list:
    """(first_row, rows) of every batch in visiting order; the last batch of the dataset may be short."""
    return [(int(b) * batch_size, min(batch_size, num_examples - int(b) * batch_size)) for b in order]
