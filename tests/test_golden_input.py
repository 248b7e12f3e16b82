"""Committed golden vectors of the input pipeline (tests/golden/input/*.json, written by tests/tools/make_golden_input.py
from oracle/input_oracle.py).  CPU: the oracle and the kernel's own source built for the host reproduce them bit
for bit.  GPU: gmvae_binarize through the C ABI reproduces the 784-wide, rank-0 cases."""
import ctypes as C
import glob
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import input_oracle as O
from tests.test_input_cpu import host_lib  # noqa: F401  (fixture: input.cuh compiled for the host)

DIR = os.path.join(os.path.dirname(__file__), "golden", "input")
CASES = sorted(glob.glob(os.path.join(DIR, "binarize_*.json")))


def _load(path):
    rec = json.load(open(path))
    inten = np.random.default_rng(rec["D"] * 7919 + rec["n_rows"]).integers(0, 256, size=(rec["n_rows"], rec["D"]), dtype=np.uint8)
    assert hashlib.sha256(inten.tobytes()).hexdigest() == rec["intensity_sha256"]
    idx = None if rec["row_index"] is None else np.asarray(rec["row_index"], dtype=np.int64)
    n = rec["batch"] * rec["D"]
    want = np.unpackbits(np.frombuffer(bytes.fromhex(rec["x_packbits_hex"]), dtype=np.uint8))[:n].reshape(rec["batch"], rec["D"])
    assert int(want.sum()) == rec["x_ones"]
    return rec, inten, idx, want


def test_golden_input_files_present():
    assert len(CASES) >= 3 and os.path.exists(os.path.join(DIR, "philox_words.json"))


def test_oracle_and_kernel_source_reproduce_philox_words(host_lib):  # noqa: F811
    for c in json.load(open(os.path.join(DIR, "philox_words.json")))["cases"]:
        got = O.philox_words(c["seed"], c["stream"], np.array([c["ctr"]], dtype=np.uint64))[0]
        assert [int(w) for w in got] == c["words"]
        out = (C.c_uint32 * 4)()
        host_lib.host_philox(c["seed"], c["stream"], c["ctr"], out)
        assert list(out) == c["words"]
        host_lib.host_philox_scheduled(c["seed"], c["stream"], c["ctr"], out)
        assert list(out) == c["words"]


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-5] for p in CASES])
def test_oracle_and_kernel_source_reproduce_golden(path, host_lib):  # noqa: F811
    rec, inten, idx, want = _load(path)
    src = np.ascontiguousarray(inten[rec["first_row"]:])
    assert (O.binarize(src, idx, rec["batch"], rec["seed"], rec["draw"], rec["rank"]) == want).all()
    out = np.zeros_like(want)
    host_lib.host_binarize(src.ctypes.data, None if idx is None else idx.ctypes.data, rec["D"], rec["batch"] * rec["D"],
                           rec["seed"], rec["draw"], rec["rank"], -1, out.ctypes.data)
    assert (out == want).all()


@pytest.mark.gpu
@pytest.mark.parametrize("path", [p for p in CASES if "_784_" in p], ids=lambda p: os.path.basename(p)[:-5])
def test_cuda_binarize_matches_golden(path):
    import torch
    import gmvae_b200
    rec, inten, idx, want = _load(path)
    assert rec["rank"] == 0 and rec["D"] == 784
    eng = gmvae_b200.Engine("vae", latent_size=10, hidden_sizes=[48], max_batch=64, seed=rec["seed"])
    dev = torch.from_numpy(inten).cuda()
    if idx is None:
        x = eng.binarize(dev, batch=rec["batch"], first_row=rec["first_row"], draw=rec["draw"])
    else:
        x = eng.binarize(dev, row_index=torch.from_numpy(idx).cuda(), draw=rec["draw"])
    assert (x.cpu().numpy() == want).all()
    eng.close()
