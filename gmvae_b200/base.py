"""Conditional distributions -- host-side mirror of /root/reference/scripts/base.py.

In the reference each class owns a Sonnet MLP and `__call__` returns a TFP distribution.  Here the variables live in the
engine's flat parameter buffer under the reference's names (`{name}_fcnet/linear_{i}/{w,b}`); `condition(tensor_list)` runs
the MLP on the caller's tensors through the C ABI (`gmvae_condition`: the same tcgen05 / SIMT GEMM kernels as the training
step) and `__call__(*tensors)` returns a small distribution object whose `sample` / `log_prob` / `mean` are the library's
own kernels (`gmvae_dist_*`, TFP semantics restated in SURVEY.md Appendix B.3-B.5).  Same constructor arguments and
defaults as the reference (base.py:18-20, 89-90, 152-153).  The fused training step (`run_model`) does not go through
these objects: it evaluates the same arithmetic inside the chained kernel."""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib

# gmvae_abi.h GMVAE_COND_*
COND_DECODER, COND_ENCODER, COND_ENCODER_Y, COND_PRIOR_GMM = 0, 1, 2, 3


def _check_activation(fn):
    if fn is None or fn == "relu":
        return
    name = getattr(fn, "__name__", str(fn))
    if name != "relu":
        raise NotImplementedError("only ReLU hidden activations are built (the reference never passes another one)")


def _f32(t, device) -> torch.Tensor:
    t = torch.as_tensor(t)
    return t.to(device=device, dtype=torch.float32).reshape(t.shape[0], -1).contiguous()


# ------------------------------------------------------------------------------------------ distribution objects
class _Distribution:
    def __init__(self, engine, name):
        self._eng, self.name = engine, name

    def _st(self):
        return self._eng._stream()

    def _noise(self, n_eps=0, n_u=0):
        """Draws from the library's device generator (Philox, keyed by the engine seed and a per-call draw counter)."""
        return self._eng.debug_noise(n_eps, n_u)


class MultivariateNormalDiag(_Distribution):
    """tfd.MultivariateNormalDiag(loc, scale_diag) as returned by ConditionalNormal.__call__ (base.py:75-83)."""

    def __init__(self, engine, loc, scale_diag, name="cond_normal"):
        super().__init__(engine, name)
        self.loc, self.scale_diag = loc, scale_diag

    def mean(self):
        return self.loc

    def stddev(self):
        return self.scale_diag

    def sample(self, eps=None):
        """loc + scale_diag * eps, eps ~ N(0, I) (drawn on the device unless injected)."""
        n = self.loc.numel()
        e = self._noise(n_eps=n)[0] if eps is None else _f32(eps, self.loc.device).reshape(-1)
        out = torch.empty_like(self.loc)
        _lib.check(self._eng.lib.gmvae_dist_normal_sample(self.loc.data_ptr(), self.scale_diag.data_ptr(), e.data_ptr(), n, out.data_ptr(),
                                                          self._st()), "gmvae_dist_normal_sample")
        self._keep = [e]
        return out

    def log_prob(self, z):
        z = _f32(z, self.loc.device)
        out = torch.empty(self.loc.shape[0], dtype=torch.float32, device=self.loc.device)
        _lib.check(self._eng.lib.gmvae_dist_normal_log_prob(self.loc.data_ptr(), self.scale_diag.data_ptr(), z.data_ptr(), self.loc.shape[0],
                                                            self.loc.shape[1], out.data_ptr(), self._st()), "gmvae_dist_normal_log_prob")
        self._keep = [z]
        return out


class IndependentBernoulli(_Distribution):
    """tfd.Independent(tfd.Bernoulli(logits), 1) as returned by ConditionalBernoulli.__call__ (base.py:138-146)."""

    def __init__(self, engine, logits, name="cond_bernoulli"):
        super().__init__(engine, name)
        self.logits = logits

    def mean(self):
        out = torch.empty_like(self.logits)
        _lib.check(self._eng.lib.gmvae_dist_bernoulli_mean(self.logits.data_ptr(), self.logits.numel(), out.data_ptr(), self._st()),
                   "gmvae_dist_bernoulli_mean")
        return out

    def log_prob(self, x):
        """sum_d x l - max(l, 0) - log1p(exp(-|l|)): minus the sigmoid cross-entropy (gmvae.py:254, vae.py:177)."""
        x = _f32(x, self.logits.device)
        out = torch.empty(self.logits.shape[0], dtype=torch.float32, device=self.logits.device)
        _lib.check(self._eng.lib.gmvae_dist_bernoulli_log_prob(self.logits.data_ptr(), x.data_ptr(), self.logits.shape[0], self.logits.shape[1],
                                                               out.data_ptr(), self._st()), "gmvae_dist_bernoulli_log_prob")
        self._keep = [x]
        return out

    def sample(self, u=None):
        p = self.mean()
        u = self._noise(n_u=p.numel())[1].reshape(p.shape) if u is None else _f32(u, p.device)
        return u < p


class _ExpRelaxed:
    """`.distribution` of TFP's RelaxedOneHotCategorical (a TransformedDistribution of ExpRelaxedOneHotCategorical): the
    reference reads the raw MLP logits through `q_y.distribution.logits` (gmvae.py:263, 271)."""

    def __init__(self, logits, temperature):
        self.logits, self.temperature = logits, temperature


class RelaxedOneHotCategorical(_Distribution):
    """tfd.RelaxedOneHotCategorical(temperature, logits) as returned by ConditionalCategorical.__call__ (base.py:201-209)."""

    def __init__(self, engine, temperature, logits, name="cond_categorical"):
        super().__init__(engine, name)
        self.temperature, self.logits = float(temperature), logits
        self.distribution = _ExpRelaxed(logits, self.temperature)

    def sample(self, u=None):
        """softmax((logits + g) / T), g = -log(-log u), u ~ U(0, 1) (drawn on the device unless injected)."""
        n, k = self.logits.shape
        uu = self._noise(n_u=n * k)[1] if u is None else _f32(u, self.logits.device).reshape(-1)
        out = torch.empty_like(self.logits)
        _lib.check(self._eng.lib.gmvae_dist_relaxed_sample(self.logits.data_ptr(), uu.data_ptr(), n, k, self.temperature, out.data_ptr(),
                                                           self._st()), "gmvae_dist_relaxed_sample")
        self._keep = [uu]
        return out


# ------------------------------------------------------------------------------------------ conditional distributions
class _Conditional:
    out_multiplier = 1
    _which: Optional[int] = None     # GMVAE_COND_* once the owning model has bound an engine

    def __init__(self, size: int, hidden_layer_sizes: Optional[List[int]], hidden_activation_fn, name: str):
        _check_activation(hidden_activation_fn)
        self._name = name
        self._size = int(size)
        self.hidden_layer_sizes = None if hidden_layer_sizes is None else [int(h) for h in hidden_layer_sizes]
        self._model = None   # set by the owning VAE/GMVAE

    @property
    def name(self) -> str:
        return self._name

    @property
    def size(self) -> int:
        return self._size

    @property
    def output_sizes(self) -> List[int]:
        """snt.nets.MLP(output_sizes=hidden + [out]) (base.py:47-60)."""
        return (self.hidden_layer_sizes or []) + [self.out_multiplier * self._size]

    def variable_names(self) -> List[str]:
        out = []
        for i in range(len(self.output_sizes)):
            out += [f"{self._name}_fcnet/linear_{i}/w", f"{self._name}_fcnet/linear_{i}/b"]
        return out

    # -- plumbing ---------------------------------------------------------------------------------
    def _bind(self, model, which: int):
        self._model, self._which = model, which

    def _run_mlp(self, tensor_list, n_out: int):
        """concat(tensor_list, axis=1) -> MLP, through gmvae_condition.  Returns (engine, out_a, out_b)."""
        if self._model is None:
            raise RuntimeError(f"{type(self).__name__} '{self._name}' is not part of a model yet: build it through create_vae / "
                               f"create_gmvae (the variables live in the model's engine)")
        tensors = list(tensor_list)
        if not tensors:
            raise ValueError("condition() needs at least one tensor")
        n = int(torch.as_tensor(tensors[0]).shape[0])
        eng = self._model.engine(n)
        dev = eng.device
        ins = [_f32(t, dev) for t in tensors]
        two = self._which == COND_ENCODER and eng.model == "gmvae"
        if two:
            if len(ins) == 1:                                   # already concatenated [x, y]
                ins = [ins[0][:, :eng.data_size].contiguous(), ins[0][:, eng.data_size:].contiguous()]
            in1, in2 = ins[0], torch.cat(ins[1:], 1) if len(ins) > 2 else ins[1]
        else:
            in1, in2 = (torch.cat(ins, 1) if len(ins) > 1 else ins[0]), None
        widths = {COND_DECODER: eng.latent_size, COND_ENCODER: eng.data_size, COND_ENCODER_Y: eng.data_size,
                  COND_PRIOR_GMM: eng.mixture_components}
        if in1.shape[1] != widths[self._which] or (two and in2.shape[1] != eng.mixture_components):
            raise ValueError(f"{self._name}: input width {in1.shape[1]}{'+' + str(in2.shape[1]) if two else ''} does not match the network")
        a = torch.empty(n, n_out, dtype=torch.float32, device=dev)
        b = torch.empty(n, n_out, dtype=torch.float32, device=dev) if self.out_multiplier == 2 else None
        for i in range(0, n, eng.max_batch):                    # the engine's buffers hold max_batch rows
            j = min(n, i + eng.max_batch)
            _lib.check(eng.lib.gmvae_condition(eng._h, self._which, in1[i:j].data_ptr(), None if in2 is None else in2[i:j].data_ptr(), j - i,
                                               a[i:j].data_ptr(), None if b is None else b[i:j].data_ptr(), eng._stream()), "gmvae_condition")
        self._keep = [in1, in2]
        return eng, a, b


class ConditionalNormal(_Conditional):
    """MultivariateNormalDiag conditioned on tensors via an MLP (base.py:15-83):
    mu, sigma = split(MLP(concat(inputs))); sigma = max(softplus(sigma + raw_sigma_bias), sigma_min)."""
    out_multiplier = 2

    def __init__(self, size, hidden_layer_sizes=None, initializers=None, sigma_min=0.0, raw_sigma_bias=0.25,
                 hidden_activation_fn="relu", name="cond_normal"):
        super().__init__(size, hidden_layer_sizes, hidden_activation_fn, name)
        self._sigma_min = float(sigma_min)
        self._raw_sigma_bias = float(raw_sigma_bias)

    def condition(self, tensor_list, **unused_kwargs):
        """(mu, sigma) of the distribution (base.py:63-72)."""
        _, mu, sigma = self._run_mlp(tensor_list, self._size)
        return mu, sigma

    def __call__(self, *args, **kwargs):
        eng, mu, sigma = self._run_mlp(args, self._size)
        return MultivariateNormalDiag(eng, mu, sigma, name=self._name)


class ConditionalBernoulli(_Conditional):
    """Independent Bernoulli with logits = MLP(inputs) + bias_init (base.py:86-146)."""

    def __init__(self, size, hidden_layer_sizes=None, initializers=None, bias_init=0.0,
                 hidden_activation_fn="relu", name="cond_bernoulli"):
        super().__init__(size, hidden_layer_sizes, hidden_activation_fn, name)
        self._bias_init = float(bias_init)

    def condition(self, tensor_list, **unused_kwargs):
        """Logits of the distribution (base.py:130-135)."""
        return self._run_mlp(tensor_list, self._size)[1]

    def __call__(self, *args, **kwargs):
        eng, logits, _ = self._run_mlp(args, self._size)
        return IndependentBernoulli(eng, logits, name=self._name)


class ConditionalCategorical(_Conditional):
    """RelaxedOneHotCategorical(temperature, logits = MLP(inputs)) (base.py:149-209)."""

    def __init__(self, size, hidden_layer_sizes=None, temperature=1.0, initializers=None,
                 hidden_activation_fn="relu", name="cond_categorical"):
        super().__init__(size, hidden_layer_sizes, hidden_activation_fn, name)
        self._temperature = float(temperature)

    def condition(self, tensor_list, **unused_kwargs):
        """Logits of the distribution (base.py:193-198)."""
        return self._run_mlp(tensor_list, self._size)[1]

    def __call__(self, *args, **kwargs):
        eng, logits, _ = self._run_mlp(args, self._size)
        return RelaxedOneHotCategorical(eng, self._temperature, logits, name=self._name)
