"""Generates tests/golden/*.json from the CPU oracle (the reference cannot be imported here:
TensorFlow 1.13 / TFP 0.6 / Sonnet are not installable, SURVEY.md F3 -- so these vectors pin the
ORACLE against regressions and give the GPU tests a committed target; they are not reference
outputs).  Inputs are regenerated from seeds; expectations are the loss terms, per-tensor gradient
norms and a few sampled gradient entries in float64.

    python tests/tools/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import gmvae_oracle as O  # noqa: E402
from tests.helpers import CONFIGS, make_spec, perturbed_params  # noqa: E402

CASES = ["tiny_vae", "tiny_gmp", "tiny_gmvae", "nohidden_gmvae", "cfg1", "cfg2", "cfg3", "run_train_sh", "cfg5_small", "k20_ragged", "k17_gmp"]


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name in CASES:
        cfg = CONFIGS[name]
        spec = make_spec(cfg)
        params = perturbed_params(spec, seed=2024)
        x, _, eps, u = O.synthetic_batch(spec, cfg["batch"])
        terms, grads = O.loss_and_grads(spec, params, x, eps, u)
        g = torch.Generator().manual_seed(99)
        samples = {}
        for n, t in grads.items():
            flat = t.reshape(-1)
            idx = torch.randint(0, flat.numel(), (4,), generator=g).tolist()
            samples[n] = [[i, flat[i].item()] for i in idx]
        rec = {
            "generator": "tests/tools/make_golden.py (oracle/gmvae_oracle.py, float64)",
            "config": cfg, "params_seed": 2024, "data_seed": 1234, "noise_seed": 4321,
            "x_sum": int(x.sum()), "eps_sum": eps.double().sum().item(),
            "terms": {k: terms[k].item() for k in ("loss", "nll", "kl_div_z", "nent")},
            "grad_norms": {n: t.norm().item() for n, t in grads.items()},
            "grad_samples": samples,
        }
        if spec.model == "gmvae":
            tm, _ = O.loss_and_grads(spec, params, x, torch.randn(cfg["batch"], spec.mixture_components, spec.latent_size,
                                                                  generator=torch.Generator().manual_seed(4321)), u, "marginal")
            rec["terms_marginal"] = {k: tm[k].item() for k in ("loss", "nll", "kl_div_z", "nent")}
        with open(os.path.join(out_dir, name + ".json"), "w") as f:
            json.dump(rec, f, indent=1)
        print(name, rec["terms"])


if __name__ == "__main__":
    main()
