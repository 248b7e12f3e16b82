"""Shared helpers of the GPU parity tests: build an Engine and an oracle Spec from one
description, copy oracle weights into the engine, compare loss terms and gradients."""
import torch

from oracle import gmvae_oracle as O

CONFIGS = {
    # BASELINE.json configs[0..2] (cfg1-3): H=512x2, Z=64, B=100
    "cfg1": dict(model="vae", latent_size=64, hidden_sizes=[512, 512], mixture_components=1, batch=100),
    "cfg2": dict(model="vae_gmp", latent_size=64, hidden_sizes=[512, 512], mixture_components=10, batch=100),
    "cfg3": dict(model="gmvae", latent_size=64, hidden_sizes=[512, 512], mixture_components=10, batch=100),
    # small / ragged shapes: nothing is a multiple of a tile
    "tiny_gmvae": dict(model="gmvae", latent_size=24, hidden_sizes=[72, 40], mixture_components=7, batch=37, data_size=200),
    "tiny_vae": dict(model="vae", latent_size=10, hidden_sizes=[48], mixture_components=1, batch=5, data_size=120),
    "tiny_gmp": dict(model="vae_gmp", latent_size=12, hidden_sizes=[40, 56], mixture_components=5, batch=33, data_size=136),
    "nohidden_gmvae": dict(model="gmvae", latent_size=16, hidden_sizes=[], mixture_components=4, batch=19, data_size=64),
    # the run_train.sh example shape (bin/run_train.sh:5-11): z=128, h=512 x 1, batch 64
    "run_train_sh": dict(model="gmvae", latent_size=128, hidden_sizes=[512], mixture_components=10, batch=64),
}


def make_spec(cfg):
    return O.Spec(model=cfg["model"], data_size=cfg.get("data_size", 784), latent_size=cfg["latent_size"],
                  hidden_sizes=list(cfg["hidden_sizes"]), mixture_components=cfg["mixture_components"])


def make_engine(cfg, precision, max_batch=None, objective="reference", **kw):
    import gmvae_b200
    return gmvae_b200.Engine(model=cfg["model"], data_size=cfg.get("data_size", 784), latent_size=cfg["latent_size"],
                             hidden_sizes=cfg["hidden_sizes"], mixture_components=cfg["mixture_components"],
                             precision=precision, objective=objective, max_batch=max_batch or cfg["batch"], init=False, **kw)


def perturbed_params(spec, seed=2024, bias_scale=0.05):
    """Xavier weights plus small non-zero biases so that every bias-gradient path is exercised."""
    p = O.init_params(spec, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    for n in p:
        if n.endswith("/b"):
            p[n] = bias_scale * torch.randn(p[n].shape, generator=g, dtype=torch.float64)
    return p


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-30)


def grad_errors(engine, grads_ref):
    """||g - g*||_2 / ||g*||_2 per tensor."""
    out = {}
    for name, g in engine.gradients().items():
        r = grads_ref[name].to(torch.float64)
        d = g.detach().cpu().to(torch.float64).reshape(r.shape)
        out[name] = ((d - r).norm() / r.norm().clamp_min(1e-30)).item()
    return out


def run_parity(cfg, precision, seed=2024, rounding_model=None, objective="reference"):
    """Returns (loss-term errors, per-tensor gradient errors) of the CUDA step against the oracle.
    With `rounding_model` the oracle restates the bf16 storage points of the CUDA path."""
    spec = make_spec(cfg)
    params = perturbed_params(spec, seed)
    x, labels, eps, u = O.synthetic_batch(spec, cfg["batch"], objective=objective)
    terms_ref, grads_ref = O.loss_and_grads(spec, params, x, eps, u, objective=objective, q=rounding_model or O.EXACT)
    eng = make_engine(cfg, precision, objective=objective)
    eng.set_parameters(params)
    loss = eng.forward_backward(x, eps=eps, gumbel_u=u)
    torch.cuda.synchronize()
    t = loss.detach().cpu().tolist()
    terr = {"loss": rel(t[0], terms_ref["loss"].item()), "nll": rel(t[1], terms_ref["nll"].item()),
            "kl_div_z": abs(t[2] - terms_ref["kl_div_z"].item()) / max(abs(terms_ref["kl_div_z"].item()), 1.0),
            "nent": abs(t[3] - terms_ref["nent"].item()) / max(abs(terms_ref["nent"].item()), 1.0)}
    gerr = grad_errors(eng, grads_ref)
    eng.close()
    return terr, gerr


# --------------------------------------------------------------------------- bf16 rounding model
class _RoundBoth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class Bf16Model(O.Exact):
    """Where the bf16 CUDA path stores reduced precision (DESIGN.md "Precision"): hidden
    activations and their gradients, z as the decoder's input, y as a GEMM operand, the gradients
    w.r.t. the decoder logits, the encoders' outputs and the prior's output, and -- when the
    tensor-core path is on -- the weight operand of every GEMM (engine.cu lin_fwd / lin_dgrad)."""

    def __init__(self, tensor_core_weights=True):
        self.tcw = tensor_core_weights

    def act(self, t):
        return _RoundBoth.apply(t)

    def fwd(self, t):
        return _RoundFwd.apply(t)

    def grad(self, t):
        return _RoundGrad.apply(t)

    def weight(self, name, w):
        return _RoundFwd.apply(w) if self.tcw else w

    def yin(self, t):            # y is the bf16 operand of the second K segment (tensor-core path)
        return _RoundFwd.apply(t) if self.tcw else t
