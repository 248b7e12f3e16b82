"""Gaussian-Mixture VAE -- mirror of /root/reference/scripts/gmvae.py."""
from __future__ import annotations

from . import base
from ._model import _EngineBacked


class GMVAE(_EngineBacked):
    """gmvae.py:11-188."""
    _model_kind = "gmvae"

    def __init__(self, mix_components, prior_gmm, decoder, encoder_y, encoder_gmm, random_seed):
        self._prior_gmm = prior_gmm
        self._decoder = decoder
        self._encoder_y = encoder_y
        self._encoder_gmm = encoder_gmm
        self.mix_components = mix_components
        self.random_seed = random_seed
        self._init_backing()
        prior_gmm._bind(self, base.COND_PRIOR_GMM); decoder._bind(self, base.COND_DECODER)
        encoder_y._bind(self, base.COND_ENCODER_Y); encoder_gmm._bind(self, base.COND_ENCODER)

    # ---- the distribution accessors of the reference (gmvae.py:49-107): each returns a distribution object ----
    def prior_gmm(self, y):
        """p(z | y): MultivariateNormalDiag [batch, latent_size] (gmvae.py:49-59)."""
        return self._prior_gmm(y)

    def decoder(self, z):
        """p(x | z): independent Bernoulli [batch, data_size] (gmvae.py:62-72)."""
        return self._decoder(z)

    def encoder_y(self, x):
        """q(y | x): RelaxedOneHotCategorical [batch, mix_components]; x is cast to float32 (gmvae.py:75-88)."""
        return self._encoder_y(x)

    def encoder_gmm(self, x, y):
        """q(z | x, y): MultivariateNormalDiag [batch, latent_size]; x is cast to float32 (gmvae.py:91-106)."""
        return self._encoder_gmm(x, y)

    def _engine_kwargs(self):
        eg, dec, ey = self._encoder_gmm, self._decoder, self._encoder_y
        if self._prior_gmm.hidden_layer_sizes is not None:
            raise NotImplementedError("prior_gmm is a single linear layer in the reference (gmvae.py:321-327)")
        if (ey.hidden_layer_sizes or []) != (eg.hidden_layer_sizes or []) or (dec.hidden_layer_sizes or []) != (eg.hidden_layer_sizes or []):
            raise NotImplementedError("all MLPs share fcnet_hidden_sizes in the reference (gmvae.py:331-353)")
        return dict(model="gmvae", data_size=dec.size, latent_size=eg.size, hidden_sizes=eg.hidden_layer_sizes or [],
                    mixture_components=self.mix_components, sigma_min=eg._sigma_min, raw_sigma_bias=eg._raw_sigma_bias,
                    gen_bias_init=dec._bias_init, temperature=ey._temperature)


    def transform(self, inputs):
        """Latent code z ~ q(z|x,y), y ~ q(y|x) (gmvae.py:140-149; the reference samples here, it does not take the mean)."""
        _, _, z = self.engine(inputs.shape[0]).encode(inputs)
        return z

    def encoder_y_logits(self, x):
        """Logits of q(y|x) (what `encoder_y(x).distribution.logits` is in the reference, gmvae.py:263,271)."""
        logits, _, _ = self.engine(x.shape[0]).encode(x)
        return logits

    def generate_samples(self, num_samples, clusters=None):
        """Samples from the prior components p(z | y = one_hot(k)) (gmvae.py:152-188): for every k (or for
        the given `clusters`) `num_samples` draws; shape [num_samples * n_clusters, latent_size], sample-major
        like the reference's reshape of [num_samples, n_clusters, latent_size]."""
        import torch
        mu, sg = self.engine().prior_table()
        if clusters is not None:
            idx = torch.as_tensor(clusters, dtype=torch.long, device=mu.device)
            mu, sg = mu[idx], sg[idx]
        eps = self._randn(num_samples, mu.shape[0], mu.shape[1]).to(mu.device)
        return (mu[None] + sg[None] * eps).reshape(num_samples * mu.shape[0], -1)


class TrainableGMVAE(GMVAE):
    """gmvae.py:191-274."""

    def __init__(self, mix_components, prior_gmm, decoder, encoder_y, encoder_gmm, random_seed=None):
        super().__init__(mix_components, prior_gmm, decoder, encoder_y, encoder_gmm, random_seed=random_seed)

    def run_model(self, images, targets, labels=None, eps=None, gumbel_u=None):
        """loss = nll + kl_div_z + nent (gmvae.py:223-274) and, in the same pass, the gradients of
        every trainable variable (runners.py:182).  `labels` only feed the cluster_acc summary in
        the reference and are not used by the step.  `eps` / `gumbel_u` inject the sampling noise."""
        return self._run(images, targets, eps, gumbel_u)


def create_gmvae(data_size, latent_size, mixture_components=1, fcnet_hidden_sizes=None, hidden_activation_fn="relu",
                 sigma_min=0.001, raw_sigma_bias=0.25, gen_bias_init=0.0, temperature=1.0, random_seed=None) -> TrainableGMVAE:
    """Factory with the reference's signature and defaults (gmvae.py:277-355)."""
    if fcnet_hidden_sizes is None:
        fcnet_hidden_sizes = [latent_size]                      # gmvae.py:316-317
    prior_gmm = base.ConditionalNormal(size=latent_size, hidden_layer_sizes=None, hidden_activation_fn=hidden_activation_fn,
                                       sigma_min=sigma_min, raw_sigma_bias=raw_sigma_bias, name="prior_gmm")
    decoder = base.ConditionalBernoulli(size=data_size, hidden_layer_sizes=fcnet_hidden_sizes,
                                        hidden_activation_fn=hidden_activation_fn, bias_init=gen_bias_init, name="decoder")
    encoder_y = base.ConditionalCategorical(size=mixture_components, temperature=temperature,
                                            hidden_layer_sizes=fcnet_hidden_sizes, hidden_activation_fn=hidden_activation_fn,
                                            name="encoder_y")
    encoder_gmm = base.ConditionalNormal(size=latent_size, hidden_layer_sizes=fcnet_hidden_sizes,
                                         hidden_activation_fn=hidden_activation_fn, sigma_min=sigma_min,
                                         raw_sigma_bias=raw_sigma_bias, name="encoder_gmm")
    return TrainableGMVAE(mixture_components, prior_gmm, decoder, encoder_y, encoder_gmm, random_seed=random_seed)
