"""The peer-memory all-reduce fused with Adam (gmvae_b200/csrc/peer.cuh): its layout and element arithmetic, compiled for the host
(tests/native/host_peer.cu) and run for `world` simulated ranks, equal a rank-ordered fp32 sum on every rank, bit for bit --
including ragged last shards and buffers shorter than the world size; and its synchronisation protocol as a host model under
ThreadSanitizer.  The CUDA kernels themselves run on 2 / 8 B200s through tools/dp_check.py (profiles/r2_dp_check_*.json)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.fixture(scope="module")
def host_peer(tmp_path_factory):
    if not (os.path.exists(NVCC) or shutil.which("nvcc")):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("host_peer") / "libhost_peer.so")
    cmd = [NVCC if os.path.exists(NVCC) else "nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
           "-Xcompiler", "-fPIC", os.path.join(ROOT, "tests", "native", "host_peer.cu"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lib = C.CDLL(out)
    lib.host_peer_allreduce.argtypes = [C.c_int, C.c_longlong, C.c_void_p]
    lib.host_peer_allreduce.restype = C.c_int
    return lib


@pytest.mark.parametrize("world,n", [(2, 8), (2, 2104700), (3, 100), (8, 2104700), (8, 12), (5, 4), (16, 6070180), (1, 16)])
def test_simulated_ranks_agree_with_rank_ordered_sum(host_peer, world, n):
    rng = np.random.default_rng(world * 1000003 + n)
    g = rng.standard_normal((world, n)).astype(np.float32)
    want = g[0].copy()
    for r in range(1, world):
        want = want + g[r]                                             # fp32, rank order: what reduce_ranks does
    buf = np.ascontiguousarray(g.copy())
    assert host_peer.host_peer_allreduce(world, n, buf.ctypes.data) == 0
    for r in range(world):
        assert (buf[r].view(np.uint32) == want.view(np.uint32)).all(), r    # identical bits on every rank


def test_rejects_bad_shapes(host_peer):
    buf = np.zeros((2, 6), dtype=np.float32)
    assert host_peer.host_peer_allreduce(2, 6, buf.ctypes.data) == -1       # not a multiple of 4 floats
    assert host_peer.host_peer_allreduce(17, 8, buf.ctypes.data) == -1


# ---------------------------------------------------------------------------- the synchronisation protocol under ThreadSanitizer
@pytest.fixture(scope="module")
def protocol_exe(tmp_path_factory):
    if not (os.path.exists(NVCC) or shutil.which("nvcc")):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("host_peer_protocol") / "host_peer_protocol")
    cmd = [NVCC if os.path.exists(NVCC) else "nvcc", "-std=c++20", "-O1", "-g", "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fsanitize=thread", "-Xcompiler", "-fno-omit-frame-pointer",
           os.path.join(ROOT, "tests", "native", "host_peer_protocol.cu"), "-o", out, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "tsan" in (r.stdout + r.stderr).lower():
        pytest.skip("ThreadSanitizer runtime not available")
    assert r.returncode == 0, r.stdout + r.stderr
    return out


def _run_protocol(exe, *args, mutate=None):
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66")
    env.pop("PEER_MUTATE", None)
    if mutate is not None:
        env["PEER_MUTATE"] = str(mutate)
    return subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, env=env, timeout=300)


@pytest.mark.parametrize("world,n,steps,blocks", [(2, 400, 40, 2), (4, 4096, 30, 3), (8, 40004, 25, 4), (5, 12, 60, 2)])
def test_protocol_is_race_free_and_exact_over_many_steps(protocol_exe, world, n, steps, blocks):
    """Epoch flags + last-block publication + single buffers reused by the next step: no data race (TSan), no
    deadlock, every rank sees the exact sums at every step -- with nothing but the flags ordering the ranks."""
    r = _run_protocol(protocol_exe, world, n, steps, blocks)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-3000:]
    assert "ThreadSanitizer" not in r.stderr
    assert '"errors": 0' in r.stdout


@pytest.mark.parametrize("mutate", [1, 2])
def test_protocol_model_detects_a_missing_wait(protocol_exe, mutate):
    """The detector works: dropping either wait (the exchange without the gradients-final flags, Adam without the landed-shard
    flags) is reported as a data race and / or wrong sums."""
    r = _run_protocol(protocol_exe, 4, 4096, 20, 3, mutate=mutate)
    assert r.returncode != 0
    assert "ThreadSanitizer: data race" in r.stderr or '"errors": 0' not in r.stdout
