"""VAE / VAE with a Gaussian-mixture prior -- mirror of /root/reference/scripts/vae.py."""
from __future__ import annotations

from typing import Optional

from . import base
from ._model import _EngineBacked


class VAE(_EngineBacked):
    """vae.py:11-123.  `prior` is None for the standard normal (vae.py:247-250) or the string
    'mixture' for the learned MixtureSameFamily prior (vae.py:231-244)."""
    _model_kind = "vae"

    def __init__(self, prior, decoder, encoder, mix_components, random_seed):
        self._prior = prior
        self._decoder = decoder
        self._encoder = encoder
        self.mix_components = mix_components
        self.random_seed = random_seed
        self._init_backing()
        decoder._bind(self, base.COND_DECODER); encoder._bind(self, base.COND_ENCODER)

    def _engine_kwargs(self):
        enc, dec = self._encoder, self._decoder
        return dict(model="vae_gmp" if self.mix_components > 1 else "vae", data_size=dec.size, latent_size=enc.size,
                    hidden_sizes=enc.hidden_layer_sizes or [], mixture_components=self.mix_components,
                    sigma_min=enc._sigma_min, raw_sigma_bias=enc._raw_sigma_bias, gen_bias_init=dec._bias_init)

    def prior(self):
        """p(z) (vae.py:41-48): the standard normal MultivariateNormalDiag(0, I) (vae.py:247-250) or the learned
        MixtureSameFamily(Categorical(mixture_logits), MVNDiag(loc, softplus(raw_scale_diag))) (vae.py:231-244)."""
        return _Prior(self)

    def decoder(self, z):
        """p(x | z): independent Bernoulli [batch, data_size] (vae.py:51-61)."""
        return self._decoder(z)

    def encoder(self, x):
        """q(z | x): MultivariateNormalDiag [batch, latent_size]; x is cast to float32 (vae.py:64-78)."""
        return self._encoder(x)

    def transform(self, inputs):
        """Mean latent code q(z|x).mean (vae.py:105-112)."""
        _, z_mean, _ = self.engine(inputs.shape[0]).encode(inputs)
        return z_mean

    def generate_samples(self, num_samples):
        """`num_samples` draws from the prior (vae.py:115-123): N(0,I), or the learned mixture
        (component ~ Categorical(mixture_logits), then N(loc_k, softplus(raw_scale_diag_k)))."""
        import torch
        eng = self.engine()
        mu, sg = eng.prior_table()
        Z = mu.shape[1]
        eps = self._randn(num_samples, Z).to(mu.device)
        if self.mix_components > 1:
            logits = eng.parameters()["mixture_logits"].detach().cpu()
            comp = torch.multinomial(torch.softmax(logits, 0), num_samples, replacement=True, generator=self._noise_gen()).to(mu.device)
            return mu[comp] + sg[comp] * eps
        return eps


class _Prior:
    """The prior object `VAE.prior()` returns: `sample(n)` and `log_prob(z)` of N(0, I) or of the learned mixture."""

    def __init__(self, model):
        self._m = model

    def sample(self, n=1):
        return self._m.generate_samples(int(n))

    def log_prob(self, z):
        import torch
        eng = self._m.engine()
        z = torch.as_tensor(z).to(eng.device, torch.float32).reshape(-1, eng.latent_size).contiguous()
        mu, sg = eng.prior_table()                              # [K, Z] (K = 1: zeros / ones)
        comps = []
        for k in range(mu.shape[0]):                            # log N(z; loc_k, s_k) through the library's kernel
            d = base.MultivariateNormalDiag(eng, mu[k:k + 1].expand(z.shape[0], -1).contiguous(), sg[k:k + 1].expand(z.shape[0], -1).contiguous())
            comps.append(d.log_prob(z))
        lp = torch.stack(comps, 1)
        if self._m.mix_components > 1:                          # MixtureSameFamily.log_prob: logsumexp over components (vae.py:240-244)
            lp = lp + torch.log_softmax(eng.parameters()["mixture_logits"], 0)[None]
        return torch.logsumexp(lp, 1)


class TrainableVAE(VAE):
    """vae.py:126-188."""

    def __init__(self, prior, decoder, encoder, mix_components=1, random_seed=None):
        super().__init__(prior, decoder, encoder, mix_components, random_seed)

    def run_model(self, images, targets, eps=None):
        """loss = nll + kl_div_z (vae.py:153-188); gradients of every variable are computed in the
        same pass (runners.py:182).  `eps` optionally injects the N(0,1) noise of q_z.sample()."""
        return self._run(images, targets, eps, None)


def create_vae(data_size, latent_size, mixture_components=1, fcnet_hidden_sizes=None, hidden_activation_fn="relu",
               sigma_min=0.001, raw_sigma_bias=0.25, gen_bias_init=0.0, random_seed=None) -> TrainableVAE:
    """Factory with the reference's signature and defaults (vae.py:191-271)."""
    if fcnet_hidden_sizes is None:
        fcnet_hidden_sizes = [latent_size]                      # vae.py:228-229
    prior = "mixture" if mixture_components > 1 else None
    decoder = base.ConditionalBernoulli(size=data_size, hidden_layer_sizes=fcnet_hidden_sizes,
                                        hidden_activation_fn=hidden_activation_fn, bias_init=gen_bias_init, name="decoder")
    encoder = base.ConditionalNormal(size=latent_size, hidden_layer_sizes=fcnet_hidden_sizes,
                                     hidden_activation_fn=hidden_activation_fn, sigma_min=sigma_min,
                                     raw_sigma_bias=raw_sigma_bias, name="encoder")
    return TrainableVAE(prior, decoder, encoder, mix_components=mixture_components, random_seed=random_seed)
