"""Generates tests/golden/input/*.json from the numpy oracle of the input pipeline (oracle/input_oracle.py).
Like the step's golden files these pin the ORACLE and give the CUDA kernel a committed target; they are not
reference outputs (TF's stateful uniform generator cannot be reproduced).  Intensities are regenerated
from a seed; the expectation is the binarised batch, bit-packed and hex-encoded.

    python tests/tools/make_golden_input.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import input_oracle as O  # noqa: E402

CASES = [
    dict(name="binarize_784_contiguous", D=784, n_rows=12, batch=6, first_row=4, gather=False, seed=0x243F6A8885A308D3, draw=17, rank=0),
    dict(name="binarize_784_gathered", D=784, n_rows=12, batch=9, first_row=0, gather=True, seed=123, draw=2 ** 40 + 5, rank=0),
    dict(name="binarize_10_ragged_rank3", D=10, n_rows=9, batch=7, first_row=0, gather=False, seed=7, draw=1, rank=3),
]


def intensities(case):
    return np.random.default_rng(case["D"] * 7919 + case["n_rows"]).integers(0, 256, size=(case["n_rows"], case["D"]), dtype=np.uint8)


def row_index(case):
    if not case["gather"]:
        return None
    return np.random.default_rng(case["batch"]).integers(0, case["n_rows"], size=case["batch"])


def expected(case):
    inten, idx = intensities(case), row_index(case)
    return O.binarize(inten[case["first_row"]:], idx, case["batch"], case["seed"], case["draw"], case["rank"])


def main():
    out_dir = os.path.join(ROOT, "tests", "golden", "input")
    os.makedirs(out_dir, exist_ok=True)
    for case in CASES:
        x = expected(case)
        rec = dict(case, generator="tests/tools/make_golden_input.py (oracle/input_oracle.py)",
                   intensity_sha256=hashlib.sha256(intensities(case).tobytes()).hexdigest(),
                   row_index=None if not case["gather"] else row_index(case).tolist(),
                   x_packbits_hex=np.packbits(x.reshape(-1)).tobytes().hex(), x_ones=int(x.sum()))
        with open(os.path.join(out_dir, case["name"] + ".json"), "w") as f:
            json.dump(rec, f, indent=1)
        print(case["name"], rec["x_ones"], "ones of", x.size)
    words = [dict(seed=s, stream=t, ctr=c, words=[int(w) for w in O.philox_words(s, t, np.array([c], dtype=np.uint64))[0]])
             for s, t, c in [(0, 0, 0), (2 ** 64 - 1, 2 ** 64 - 1, 2 ** 64 - 1), (0x299F31D0A4093822, 0x0370734413198A2E, 0x85A308D3243F6A88),
                             (1234, O.BINARIZE_STREAM, 0), (1234, O.BINARIZE_STREAM + 1, 2 ** 33 + 9)]]
    with open(os.path.join(out_dir, "philox_words.json"), "w") as f:
        json.dump({"generator": "tests/tools/make_golden_input.py", "note": "first three = Random123 kat_vectors (philox4x32, 10 rounds)",
                   "cases": words}, f, indent=1)


if __name__ == "__main__":
    main()
