"""Consumes the golden vectors tools/dump_reference.py writes from the UNMODIFIED reference (TF 1.13 / TFP 0.6 / Sonnet v1).
With a dump present under tests/golden/reference/ the oracle (CPU) and the CUDA step (GPU, fp32 validation mode) must reproduce
the reference's loss terms, every gradient and one Adam step; with none (TF cannot be installed in this image) the tests skip and
say so.  The loader itself is exercised on every run through a dump of the same layout written from the oracle."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import gmvae_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
DUMPS = sorted(glob.glob(os.path.join(HERE, "golden", "reference", "*.npz")))
TOL = 1e-5          # the reference computes in fp32


def load_dump(path):
    z = np.load(path, allow_pickle=False)
    cfg = json.loads(str(z["config"]))
    spec = O.Spec(model=cfg["model"], data_size=784, latent_size=cfg["latent_size"], hidden_sizes=[cfg["hidden_size"]] * cfg["num_layers"],
                  mixture_components=cfg["mixture_components"])
    params = {k[len("param/"):]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("param/")}
    assert set(params) == {n for n, _ in O.param_shapes(spec)}, "variable names differ from the reference's"
    grads = {k[len("grad/"):]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("grad/")}
    adam1 = {k[len("adam1/"):]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("adam1/")}
    terms = {k: float(z["term/" + k]) for k in ("loss", "nll", "kl_div_z", "nent")}
    x = torch.from_numpy(z["x"]).bool()
    eps = torch.from_numpy(z["eps"])
    u = torch.from_numpy(z["u"]) if "u" in z.files else None
    return cfg, spec, params, x, eps, u, terms, grads, adam1


def write_dump_from_oracle(path, cfg):
    """A dump of the SAME layout as tools/dump_reference.py's, produced by the oracle (exercises the loader; pins nothing)."""
    spec = O.Spec(model=cfg["model"], data_size=784, latent_size=cfg["latent_size"], hidden_sizes=[cfg["hidden_size"]] * cfg["num_layers"],
                  mixture_components=cfg["mixture_components"])
    params = O.init_params(spec, seed=cfg["seed"])
    x, _, eps, u = O.synthetic_batch(spec, cfg["batch_size"])
    terms, grads = O.loss_and_grads(spec, params, x, eps, u)
    out = {"config": json.dumps(cfg), "x": x.numpy().astype(np.uint8), "eps": eps.numpy()}
    if u is not None:
        out["u"] = u.numpy()
    for k in ("loss", "nll", "kl_div_z", "nent"):
        out["term/" + k] = np.float64(terms[k].item())
    for n, g in grads.items():
        out["grad/" + n] = g.numpy().astype(np.float32)
        out["param/" + n] = params[n].numpy().astype(np.float32)
    p1 = {k: v.float().double() for k, v in params.items()}
    O.adam_tf_step(p1, {k: v for k, v in grads.items()}, O.adam_init(p1), lr=cfg["learning_rate"])
    for n, v in p1.items():
        out["adam1/" + n] = v.numpy().astype(np.float32)
    np.savez_compressed(path, **out)


def check_oracle_against(path):
    cfg, spec, params, x, eps, u, terms, grads, adam1 = load_dump(path)
    t, g = O.loss_and_grads(spec, params, x, eps, u)
    for k, v in terms.items():
        assert abs(t[k].item() - v) <= TOL * max(abs(v), 1e-30) or (v == 0.0 and t[k].item() == 0.0), (k, t[k].item(), v)
    for n, ref in grads.items():
        assert ((g[n] - ref.reshape(g[n].shape)).norm() / ref.norm().clamp_min(1e-30)).item() < TOL, n
    p1 = {k: v.clone() for k, v in params.items()}
    O.adam_tf_step(p1, g, O.adam_init(p1), lr=cfg["learning_rate"])
    for n, ref in adam1.items():
        assert ((p1[n] - ref.reshape(p1[n].shape)).norm() / ref.norm().clamp_min(1e-30)).item() < TOL, ("adam", n)


def test_loader_roundtrip(tmp_path):
    cfg = dict(model="gmvae", latent_size=6, hidden_size=24, num_layers=2, mixture_components=5, batch_size=9, learning_rate=1e-3, seed=7)
    path = str(tmp_path / "roundtrip.npz")
    write_dump_from_oracle(path, cfg)
    check_oracle_against(path)


@pytest.mark.skipif(not DUMPS, reason="no reference dump under tests/golden/reference/ (TF 1.13 is not installable here; "
                                      "run tools/dump_reference.py where it is) -- the oracle stays 'parity unpinned'")
@pytest.mark.parametrize("path", DUMPS or ["none"], ids=[os.path.basename(p) for p in DUMPS] or ["none"])
def test_oracle_matches_reference_dump(path):
    check_oracle_against(path)


@pytest.mark.gpu
@pytest.mark.skipif(not DUMPS, reason="no reference dump under tests/golden/reference/")
@pytest.mark.parametrize("path", DUMPS or ["none"], ids=[os.path.basename(p) for p in DUMPS] or ["none"])
def test_cuda_step_matches_reference_dump(path):
    import gmvae_b200
    cfg, spec, params, x, eps, u, terms, grads, adam1 = load_dump(path)
    eng = gmvae_b200.Engine(model=cfg["model"], data_size=784, latent_size=cfg["latent_size"], hidden_sizes=spec.hidden_sizes,
                            mixture_components=cfg["mixture_components"], precision="fp32", max_batch=cfg["batch_size"],
                            learning_rate=cfg["learning_rate"], init=False)
    eng.set_parameters(params)
    t = eng.forward_backward(x, eps=eps, gumbel_u=u).cpu().tolist()     # (train_step clears the gradient buffer once Adam has read it)
    for i, k in enumerate(("loss", "nll", "kl_div_z", "nent")):
        assert abs(t[i] - terms[k]) <= TOL * max(abs(terms[k]), 1e-30) or (terms[k] == 0.0 and t[i] == 0.0), (k, t[i], terms[k])
    for n, gv in eng.gradients().items():
        ref = grads[n]
        assert ((gv.cpu().double().reshape(ref.shape) - ref).norm() / ref.norm().clamp_min(1e-30)).item() < TOL, n
    eng.adam_step()
    for n, pv in eng.parameters().items():
        ref = adam1[n]
        assert ((pv.cpu().double().reshape(ref.shape) - ref).norm() / ref.norm().clamp_min(1e-30)).item() < TOL, ("adam", n)
    eng.close()
