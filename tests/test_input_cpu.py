"""Input pipeline, CPU side: the numpy oracle against published known answers, and
the CUDA kernel's own `__host__ __device__` source (gmvae_b200/csrc/input.cuh, compiled for the host by
tests/native/host_input.cu) against the oracle, bit for bit.  The GPU run of the same kernel is
tests/test_input_gpu.py."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import input_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

# Random123 (D. E. Shaw Research) kat_vectors, philox4x32 10 rounds: counter, key -> output
PHILOX_KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF),
     (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_oracle_philox_known_answers():
    for ctr, key, want in PHILOX_KAT:
        got = O.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert tuple(int(x) for x in got) == want


def test_oracle_u01_open_interval():
    u = O.u01(np.array([0, 0x1FF, 0x200, 0xFFFFFFFF], dtype=np.uint32))
    assert u.dtype == np.float32
    assert u[0] == np.float32(2.0 ** -24) and u[1] == u[0] and u[2] == np.float32(1.5 * 2.0 ** -23)
    assert u[3] == np.float32(1.0 - 2.0 ** -24) and u[3] < 1.0


def test_oracle_binarize_edges_and_distribution():
    # runners.py:44-47, inverted: intensity 255 -> 1.0 < u never holds -> 0; intensity 0 -> 0.0 < u always holds -> 1
    inten = np.zeros((64, 784), dtype=np.uint8)
    inten[1::2] = 255
    x = O.binarize(inten, None, 64, seed=7, draw=3)
    assert x.dtype == np.uint8 and x.shape == (64, 784)
    assert (x[0::2] == 1).all() and (x[1::2] == 0).all()
    # P(x = 1) = 1 - v/255
    for v in (32, 128, 200):
        inten = np.full((512, 784), v, dtype=np.uint8)
        p = O.binarize(inten, None, 512, seed=11, draw=v).mean()
        assert abs(p - (1.0 - v / 255.0)) < 4e-3, (v, p)
    # a new draw counter, seed or rank gives new uniforms; the same triple repeats them
    inten = np.full((8, 784), 128, dtype=np.uint8)
    a = O.binarize(inten, None, 8, seed=1, draw=0)
    assert (a == O.binarize(inten, None, 8, seed=1, draw=0)).all()
    for kw in (dict(seed=2, draw=0), dict(seed=1, draw=1), dict(seed=1, draw=0, rank=1)):
        assert (a != O.binarize(inten, None, 8, **kw)).mean() > 0.3


def test_oracle_batch_order_is_batch_level():
    # runners.py:50-57: batch first, shuffle after -> batches are contiguous runs, the short tail batch exists
    order = O.batch_order(1050, 100, np.random.default_rng(0), shuffle=True)
    assert sorted(order.tolist()) == list(range(11))
    b = O.epoch_batches(1050, 100, order)
    assert sorted(b)[-1] == (1000, 50) and sum(n for _, n in b) == 1050
    assert O.epoch_batches(1050, 100, O.batch_order(1050, 100, None, shuffle=False))[0] == (0, 100)


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    if not (os.path.exists(NVCC) or shutil.which("nvcc")):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("host_input") / "libhost_input.so")
    cmd = [NVCC if os.path.exists(NVCC) else "nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
           "-Xcompiler", "-fPIC", os.path.join(ROOT, "tests", "native", "host_input.cu"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lib = C.CDLL(out)
    lib.host_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
    lib.host_philox.restype = None
    lib.host_u01.argtypes = [C.c_uint32]
    lib.host_u01.restype = C.c_float
    lib.host_binarize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    lib.host_binarize.restype = C.c_int
    lib.host_threshold_agrees.argtypes = [C.c_uint32, C.c_uint32]
    lib.host_threshold_agrees.restype = C.c_int
    lib.host_threshold.argtypes = [C.c_uint32]
    lib.host_threshold.restype = C.c_uint32
    lib.host_philox_scheduled.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
    lib.host_philox_scheduled.restype = None
    return lib


def test_kernel_source_philox_known_answers(host_lib):
    """The library's Philox struct (the one the step's noise and the binarisation use) reproduces the
    Random123 vectors: counter = (ctr_lo, ctr_hi, stream_lo, stream_hi), key = (seed_lo, seed_hi)."""
    for ctr, key, want in PHILOX_KAT:
        out = (C.c_uint32 * 4)()
        host_lib.host_philox(key[0] | (key[1] << 32), ctr[2] | (ctr[3] << 32), ctr[0] | (ctr[1] << 32), out)
        assert tuple(out) == want
    for r in (0, 0x1FF, 0x200, 0x12345678, 0xFFFFFFFF):
        assert np.float32(host_lib.host_u01(r)) == O.u01(np.array([r], dtype=np.uint32))[0]


def test_kernel_source_scheduled_philox_and_threshold_table(host_lib):
    """The kernel's two shortcuts are exact: Philox with the key schedule expanded on the host equals Philox::gen, and
    `(r >> 9) >= T[v]` equals the reference's `float32(v) / 255 < uniform` (runners.py:45-46) -- checked at the two
    uniforms around every threshold, at the extremes and at random words."""
    for ctr, key, want in PHILOX_KAT:
        out = (C.c_uint32 * 4)()
        host_lib.host_philox_scheduled(key[0] | (key[1] << 32), ctr[2] | (ctr[3] << 32), ctr[0] | (ctr[1] << 32), out)
        assert tuple(out) == want
    rng = np.random.default_rng(0)
    for _ in range(200):
        seed, stream, ctr = (int(x) for x in rng.integers(0, 2 ** 63, size=3))
        a, b = (C.c_uint32 * 4)(), (C.c_uint32 * 4)()
        host_lib.host_philox(seed, stream, ctr, a)
        host_lib.host_philox_scheduled(seed, stream, ctr, b)
        assert tuple(a) == tuple(b)
    assert host_lib.host_threshold(0) == 0 and host_lib.host_threshold(255) == 2 ** 23
    words = [int(x) for x in rng.integers(0, 2 ** 32, size=64)]
    for v in range(256):
        t = host_lib.host_threshold(v)
        assert 0 <= t <= 2 ** 23
        probe = [0, 0xFFFFFFFF] + words
        for m in (t - 2, t - 1, t, t + 1):
            if 0 <= m < 2 ** 23:
                probe += [m << 9, (m << 9) | 0x1FF]
        assert all(host_lib.host_threshold_agrees(v, r) for r in probe), v
    # the thresholds are the oracle's: x = 1 exactly for the uniforms above fl(v / 255)
    v = np.arange(256, dtype=np.uint8)
    unit = v.astype(np.float32) / np.float32(255.0)
    T = np.array([host_lib.host_threshold(int(i)) for i in v], dtype=np.int64)
    below = O.u01(((np.maximum(T - 1, 0)) << 9).astype(np.uint32))
    assert (unit[T > 0] >= below[T > 0]).all()
    at = O.u01((np.minimum(T, 2 ** 23 - 1) << 9).astype(np.uint32))
    assert (unit[T < 2 ** 23] < at[T < 2 ** 23]).all()


def _run_host(lib, inten, row_index, batch, seed, draw, rank, mode):
    D = inten.shape[1]
    out = np.full((batch, D), 0xEE, dtype=np.uint8)
    idx = None if row_index is None else np.ascontiguousarray(row_index, dtype=np.int64)
    used = lib.host_binarize(inten.ctypes.data, None if idx is None else idx.ctypes.data, D, batch * D, seed, draw, rank, int(mode),
                             out.ctypes.data)
    assert used == mode or mode < 0
    return out


BYTES, VEC4, VEC16 = 0, 1, 2


@pytest.mark.parametrize("D,batch,mode", [(784, 33, VEC16), (784, 33, VEC4), (784, 33, BYTES), (10, 7, BYTES), (3, 5, BYTES),
                                          (8, 1, VEC4), (16, 1, VEC16), (48, 257, VEC16), (20, 19, VEC4)])
def test_kernel_source_matches_oracle(host_lib, D, batch, mode):
    rng = np.random.default_rng(D * 1000 + batch)
    n_rows = batch + 9
    inten = rng.integers(0, 256, size=(n_rows, D), dtype=np.uint8)
    seed, draw, rank = 0x243F6A8885A308D3, 12345, 3
    got = _run_host(host_lib, inten, None, batch, seed, draw, rank, mode)
    assert (got == O.binarize(inten, None, batch, seed, draw, rank)).all()
    idx = rng.integers(0, n_rows, size=batch)                              # gathered rows, repeats allowed
    got = _run_host(host_lib, inten, idx, batch, seed, draw, rank, mode)
    assert (got == O.binarize(inten, idx, batch, seed, draw, rank)).all()
    assert set(np.unique(got)) <= {0, 1}


def test_kernel_source_all_intensities_and_mode_choice(host_lib):
    # every byte value against many uniforms, on every path
    inten = np.tile(np.arange(256, dtype=np.uint8), (64, 1))               # D = 256
    want = O.binarize(inten, None, 64, 99, 1, 0)
    for mode in (BYTES, VEC4, VEC16):
        got = _run_host(host_lib, inten, None, 64, 99, 1, 0, mode)
        assert (got == want).all()
        assert (got[:, 0] == 1).all() and (got[:, 255] == 0).all()
    # the widest path the shapes and alignments allow is the one chosen
    buf = np.zeros(64 * 48 + 64, dtype=np.uint8)
    base = (-buf.ctypes.data) % 16
    out = np.zeros(64 * 48 + 64, dtype=np.uint8)
    obase = (-out.ctypes.data) % 16
    for off, D, want_mode in ((0, 48, VEC16), (4, 48, VEC4), (1, 48, BYTES), (0, 20, VEC4), (0, 10, BYTES)):
        src = buf[base + off:]
        used = host_lib.host_binarize(src.ctypes.data, None, D, 4 * D, 1, 2, 0, -1, out[obase:].ctypes.data)
        assert used == want_mode, (off, D, used)
