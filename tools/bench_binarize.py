"""Times gmvae_binarize (csrc/input.cuh) alone on cuda:0 with CUDA events: algorithmic bytes = 2 * D per sample
(D intensity bytes in, D binarised bytes out), against the measured HBM copy bandwidth of MEASURED_PEAKS.json.
Two sizes: one cfg4 batch (16 384 rows, L2-resident on a second visit -- flushed between launches) and 400 000 rows
(627 MB of traffic per launch: larger than the 126 MB L2).  Prints one JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmvae_b200  # noqa: E402


def time_launches(fn, iters, flush=None):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        if flush is not None:
            flush.add_(1)
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2], t[0]


def main():
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    eng = gmvae_b200.Engine("vae", latent_size=8, hidden_sizes=[64], max_batch=64, seed=1)
    D = 784
    out = {"kernel": "binarize_kernel", "peak_hbm_gbs": peaks["hbm_gbs"], "cases": []}
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")          # 256 MB
    for rows, use_flush in ((16384, True), (400000, False)):
        inten = torch.randint(0, 256, (rows, D), dtype=torch.uint8, device="cuda")
        x = torch.empty(rows, D, dtype=torch.uint8, device="cuda")
        draw = [0]

        def fn():
            eng.binarize(inten, out=x, draw=draw[0]); draw[0] += 1
        for _ in range(5):
            fn()
        med, best = time_launches(fn, 30, flush if use_flush else None)
        gb = 2.0 * rows * D / 1e9
        out["cases"].append({"rows": rows, "ms_median": med, "ms_best": best, "achieved_gbs": gb / (med * 1e-3),
                             "frac_of_hbm_peak": gb / (med * 1e-3) / peaks["hbm_gbs"], "samples_per_s": rows / (med * 1e-3),
                             "l2": "flushed (256 MB write) before each launch" if use_flush else "traffic per launch exceeds L2"})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
