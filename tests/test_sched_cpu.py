"""The chained kernel's schedule arithmetic (gmvae_b200/csrc/chain_sched.cuh: chain_job_geometry on the host, chain_walk / chain_tile on
the device) compiled for the host (tests/native/host_chain_sched.cu) and simulated over every walker and CTA: every tile of a job's
(k-split x row block x n-tile) space is taken exactly once -- single CTAs, CTA pairs (phantom halves), 4-CTA clusters (double tiles,
phantom pair tiles, shared A rows), with and without the walker partition of the backward pass, for the grid sizes a launch can get."""
import ctypes as C
import itertools
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    if not (os.path.exists(NVCC) or shutil.which("nvcc")):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("host_sched") / "libhost_sched.so")
    cmd = [NVCC if os.path.exists(NVCC) else "nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
           "-Xcompiler", "-fPIC", os.path.join(ROOT, "tests", "native", "host_chain_sched.cu"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    L = C.CDLL(out)
    L.host_sched_check_gemm.argtypes = [C.c_int] * 11
    L.host_sched_check_rows.argtypes = [C.c_int] * 6
    return L


# (M, N, block_n, requested k-splits, k-blocks, a_mn): the jobs of cfg4 / cfg5 / the ragged test shapes
JOBS = [
    (16384, 512, 256, 1, 13, 0), (16384, 784, 256, 1, 8, 0), (16384, 10, 16, 1, 8, 0), (16384, 128, 128, 1, 8, 0),
    (512, 784, 256, 9, 256, 1), (784, 512, 256, 9, 256, 1), (10, 512, 256, 37, 256, 1), (512, 10, 16, 37, 256, 1), (64, 512, 256, 37, 256, 1),
    (65536, 1024, 256, 1, 16, 0), (1024, 1024, 256, 4, 1024, 1),
    (5000, 512, 256, 1, 8, 0), (4200, 784, 256, 1, 8, 0), (100, 512, 256, 1, 13, 0), (37, 72, 128, 1, 4, 0), (150, 96, 128, 1, 2, 0),
    (130, 640, 256, 1, 3, 0),          # three n-tiles: no shared rows in quad mode, phantom pair tile
    (300, 300, 256, 7, 5, 1), (1, 1, 16, 1, 1, 0),
]


@pytest.mark.parametrize("cl", [1, 2, 4])
def test_every_tile_exactly_once(lib, cl):
    for (M, N, bn, split, kb, a_mn), G, base in itertools.product(JOBS, [1, 3, 33, 37, 74, 148], [0, 5, 123]):
        assert lib.host_sched_check_gemm(M, N, bn, split, kb, a_mn, cl, G, base, 0, 0) == 0, (cl, M, N, bn, split, kb, a_mn, G, base)


@pytest.mark.parametrize("cl", [1, 2, 4])
def test_walker_partition(lib, cl):
    """Chain jobs on walkers [0, x), weight gradients on [x, W); a grid smaller than planned falls back to every walker."""
    for (M, N, bn, split, kb, a_mn), (W, x), base in itertools.product(JOBS, [(74, 44), (74, 12), (37, 20), (8, 4)], [0, 17]):
        for wf, wc in ((0, x), (x, W - x)):
            assert lib.host_sched_check_gemm(M, N, bn, split, kb, a_mn, cl, W, base, wf, wc) == 0, (cl, M, N, W, x, wf, wc)
            assert lib.host_sched_check_gemm(M, N, bn, split, kb, a_mn, cl, max(1, x - 1), base, wf, wc) == 0   # smaller grid


@pytest.mark.parametrize("cl", [1, 2, 4])
def test_row_jobs(lib, cl):
    for total, G, base, (wf, wc) in itertools.product([1, 7, 128, 512, 4096], [1, 5, 37, 74], [0, 3, 296], [(0, 0), (0, 3), (3, 2)]):
        if wc and wf + wc > G:
            continue
        assert lib.host_sched_check_rows(total, cl, G, base, wf, wc) == 0, (cl, total, G, base, wf, wc)
