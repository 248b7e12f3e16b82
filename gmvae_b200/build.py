"""In-tree build of libgmvae_b200.so (nvcc, sm_100a only).  `python -m gmvae_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgmvae_b200.so")
SOURCES = ["engine.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "gmvae_abi.h")]


def nccl_dirs():
    site = sysconfig.get_paths()["purelib"]
    base = os.path.join(site, "nvidia", "nccl")
    return os.path.join(base, "include"), os.path.join(base, "lib")


HASH = LIB + ".srchash"


def source_hash() -> str:
    """sha256 over the sources and headers the library is built from (content, not mtime: the snapshot that carries the
    built library to the GPU box does not keep timestamps in order)."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(os.path.join(CSRC, s) for s in SOURCES + HEADERS):
        if os.path.exists(f):
            h.update(os.path.basename(f).encode())
            with open(f, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    try:
        with open(HASH) as fh:
            return fh.read().strip() != source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    inc, libdir = nccl_dirs()
    cmd = [nvcc, "-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
           "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
           "-I", inc, "-I", os.path.join(HERE, "..", "include")]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += os.environ.get("GMVAE_NVCC_FLAGS", "").split()      # experiments (e.g. -DGMVAE_PAIR_STAGES=4)
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-o", LIB, "-L", libdir, "-l:libnccl.so.2", "-Xlinker", "-rpath", "-Xlinker", libdir, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libgmvae_b200.so")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    with open(HASH, "w") as fh:
        fh.write(source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
