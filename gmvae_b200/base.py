"""Conditional distributions -- host-side mirror of /root/reference/scripts/base.py.

In the reference each class owns a Sonnet MLP and returns a TFP distribution.  Here each object is
a *description* of one MLP head (sizes, hyper-parameters, variable-name scope); the variables live
in the engine's flat parameter buffer under the reference's names (`{name}_fcnet/linear_{i}/{w,b}`),
and the arithmetic of `condition()` runs inside the fused CUDA step.  Same constructor arguments,
same defaults (base.py:18-20, 89-90, 152-153)."""
from __future__ import annotations

from typing import List, Optional


def _check_activation(fn):
    if fn is None or fn == "relu":
        return
    name = getattr(fn, "__name__", str(fn))
    if name != "relu":
        raise NotImplementedError("only ReLU hidden activations are built (the reference never passes another one)")


class _Conditional:
    out_multiplier = 1

    def __init__(self, size: int, hidden_layer_sizes: Optional[List[int]], hidden_activation_fn, name: str):
        _check_activation(hidden_activation_fn)
        self._name = name
        self._size = int(size)
        self.hidden_layer_sizes = None if hidden_layer_sizes is None else [int(h) for h in hidden_layer_sizes]
        self._model = None   # set by the owning VAE/GMVAE once an engine exists

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


class ConditionalNormal(_Conditional):
    """MultivariateNormalDiag conditioned on tensors via an MLP (base.py:15-83):
    mu, sigma = split(MLP(concat(inputs))); sigma = max(softplus(sigma + raw_sigma_bias), sigma_min)."""
    out_multiplier = 2

    def __init__(self, size, hidden_layer_sizes=None, initializers=None, sigma_min=0.0, raw_sigma_bias=0.25,
                 hidden_activation_fn="relu", name="cond_normal"):
        super().__init__(size, hidden_layer_sizes, hidden_activation_fn, name)
        self._sigma_min = float(sigma_min)
        self._raw_sigma_bias = float(raw_sigma_bias)


class ConditionalBernoulli(_Conditional):
    """Independent Bernoulli with logits = MLP(inputs) + bias_init (base.py:86-146)."""

    def __init__(self, size, hidden_layer_sizes=None, initializers=None, bias_init=0.0,
                 hidden_activation_fn="relu", name="cond_bernoulli"):
        super().__init__(size, hidden_layer_sizes, hidden_activation_fn, name)
        self._bias_init = float(bias_init)


class ConditionalCategorical(_Conditional):
    """RelaxedOneHotCategorical(temperature, logits = MLP(inputs)) (base.py:149-209)."""

    def __init__(self, size, hidden_layer_sizes=None, temperature=1.0, initializers=None,
                 hidden_activation_fn="relu", name="cond_categorical"):
        super().__init__(size, hidden_layer_sizes, hidden_activation_fn, name)
        self._temperature = float(temperature)
