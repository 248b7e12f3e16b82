"""The callable distribution layer (base.py:15-209) and the model accessors (gmvae.py:49-107, vae.py:41-78): `condition()`,
`__call__()` and the returned distribution objects against the oracle's blocks, through the C ABI (gmvae_condition,
gmvae_dist_*).  The last test writes the reference's run_model body (gmvae.py:238-267) with the accessors, line for line, and
compares the loss with the oracle and with the fused step."""
import pytest
import torch

from oracle import gmvae_oracle as O
from tests.helpers import CONFIGS, make_spec, perturbed_params

pytestmark = pytest.mark.gpu


def _gmvae(precision="fp32", cfg_name="cfg3"):
    import gmvae_b200
    cfg = CONFIGS[cfg_name]
    spec = make_spec(cfg)
    m = gmvae_b200.create_gmvae(spec.data_size, spec.latent_size, mixture_components=spec.mixture_components,
                                fcnet_hidden_sizes=list(spec.hidden_sizes), sigma_min=0.0, raw_sigma_bias=0.5, random_seed=11)
    m.configure(precision=precision, max_batch=cfg["batch"])
    params = perturbed_params(spec)
    m.engine().set_parameters(params)
    x, _, eps, u = O.synthetic_batch(spec, cfg["batch"])
    return m, spec, params, x, eps, u


def _close(a, b, tol):
    b = b.to(torch.float64)
    return ((a.detach().cpu().double() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item() < tol


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_conditional_normal(precision, tol):
    """ConditionalNormal.condition / __call__ (base.py:63-83): encoder_gmm(x, y) and prior_gmm(y)."""
    m, spec, params, x, eps, u = _gmvae(precision)
    L = len(spec.hidden_sizes) + 1
    y = torch.softmax(torch.randn(x.shape[0], spec.K, generator=torch.Generator().manual_seed(3), dtype=torch.float64), 1)
    mu_ref, sg_ref = O.normal_params(spec, O.mlp(params, "encoder_gmm", torch.cat([x.double(), y], 1), L))
    mu, sg = m._encoder_gmm.condition([x, y])
    assert _close(mu, mu_ref, tol) and _close(sg, sg_ref, tol)
    q = m.encoder_gmm(x, y)                                           # the accessor (gmvae.py:91-106) -> distribution object
    assert _close(q.mean(), mu_ref, tol)
    z = q.sample(eps)
    assert _close(z, q.loc.cpu().double() + q.scale_diag.cpu().double() * eps.double(), 1e-6)
    lp_ref = O.mvn_diag_log_prob(z.cpu().double(), q.loc.cpu().double(), q.scale_diag.cpu().double())
    assert _close(q.log_prob(z), lp_ref, 1e-5)
    pm_ref, ps_ref = O.normal_params(spec, O.mlp(params, "prior_gmm", y, 1))
    p = m.prior_gmm(y)
    assert _close(p.loc, pm_ref, tol) and _close(p.scale_diag, ps_ref, tol)
    z2 = q.sample()                                                  # device noise: fresh on every call, standard normal
    z3 = q.sample()
    assert not torch.equal(z2, z3)
    e = (z2 - q.loc) / q.scale_diag
    assert abs(e.mean().item()) < 0.05 and abs(e.std().item() - 1.0) < 0.05


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_conditional_bernoulli(precision, tol):
    """ConditionalBernoulli.condition / __call__ (base.py:130-146): decoder(z)."""
    m, spec, params, x, eps, u = _gmvae(precision)
    L = len(spec.hidden_sizes) + 1
    z = torch.randn(x.shape[0], spec.latent_size, generator=torch.Generator().manual_seed(5))
    logits_ref = O.mlp(params, "decoder", z.double(), L) + spec.gen_bias_init
    assert _close(m._decoder.condition([z]), logits_ref, tol)
    p = m.decoder(z)
    assert _close(p.logits, logits_ref, tol)
    lp_ref = O.bernoulli_log_prob(x.double(), p.logits.cpu().double())
    assert _close(p.log_prob(x), lp_ref, 1e-5)
    assert _close(p.mean(), torch.sigmoid(p.logits.cpu().double()), 1e-5)
    s = p.sample()
    assert s.dtype == torch.bool and tuple(s.shape) == tuple(x.shape)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_conditional_categorical(precision, tol):
    """ConditionalCategorical.condition / __call__ (base.py:193-209): encoder_y(x)."""
    m, spec, params, x, eps, u = _gmvae(precision)
    L = len(spec.hidden_sizes) + 1
    logits_ref = O.mlp(params, "encoder_y", x.double(), L)
    assert _close(m._encoder_y.condition([x]), logits_ref, tol)
    q = m.encoder_y(x)
    assert _close(q.distribution.logits, logits_ref, tol) and q.temperature == 1.0
    y = q.sample(u)
    y_ref = O.gumbel_softmax_sample(q.logits.cpu().double(), u.double(), 1.0)
    assert _close(y, y_ref, 1e-5)
    assert (y.sum(1) - 1).abs().max().item() < 1e-5
    assert not torch.equal(q.sample(), q.sample())


def test_vae_accessors_and_priors():
    """VAE.encoder / decoder / prior (vae.py:41-78), standard-normal and mixture priors (vae.py:231-250)."""
    import gmvae_b200
    for name in ("cfg1", "cfg2"):
        cfg = CONFIGS[name]
        spec = make_spec(cfg)
        m = gmvae_b200.create_vae(spec.data_size, spec.latent_size, mixture_components=spec.mixture_components,
                                  fcnet_hidden_sizes=list(spec.hidden_sizes), sigma_min=0.0, raw_sigma_bias=0.5, random_seed=3)
        m.configure(precision="fp32", max_batch=cfg["batch"])
        params = perturbed_params(spec)
        m.engine().set_parameters(params)
        x, _, eps, _ = O.synthetic_batch(spec, cfg["batch"])
        L = len(spec.hidden_sizes) + 1
        mu_ref, sg_ref = O.normal_params(spec, O.mlp(params, "encoder", x.double(), L))
        q = m.encoder(x)
        assert _close(q.loc, mu_ref, 1e-5) and _close(q.scale_diag, sg_ref, 1e-5)
        z = q.sample(eps)
        lp = m.prior().log_prob(z)
        zd = z.cpu().double()
        if name == "cfg1":
            ref = -0.5 * (zd ** 2).sum(-1) - 0.5 * spec.latent_size * O.LOG_2PI
        else:
            comp = O.mvn_diag_log_prob(zd[:, None, :], params["loc"][None], O.softplus(params["raw_scale_diag"])[None])
            ref = torch.logsumexp(comp + torch.log_softmax(params["mixture_logits"], -1)[None], dim=-1)
        assert _close(lp, ref, 1e-5)
        assert tuple(m.prior().sample(7).shape) == (7, spec.latent_size)
        assert _close(m.decoder(z).logits, O.mlp(params, "decoder", zd, L), 1e-5)


def test_run_model_composed_from_accessors():
    """The reference's run_model body (gmvae.py:238-267) written with the accessors, statement for statement, gives the
    oracle's loss terms and the fused step's."""
    m, spec, params, x, eps, u = _gmvae("fp32")
    q_y = m.encoder_y(x)                                              # gmvae.py:238
    y = q_y.sample(u)                                                 # :240
    p_z_given_y = m.prior_gmm(y)                                      # :243
    q_z = m.encoder_gmm(x, y)                                         # :246
    z = q_z.sample(eps)                                               # :248
    p_x_given_z = m.decoder(z)                                        # :251
    nll = -p_x_given_z.log_prob(x).mean()                             # :254
    kl_div_z = (q_z.log_prob(z) - p_z_given_y.log_prob(z)).mean()     # :258
    logits = q_y.distribution.logits                                  # :263
    nent = (torch.softmax(logits, 1) * torch.log_softmax(logits, 1)).sum(1).mean()   # utils.entropy, :262
    loss = nll + kl_div_z + nent                                      # :267
    ref = O.loss_terms(spec, params, x, eps, u)
    for got, k in ((loss, "loss"), (nll, "nll"), (kl_div_z, "kl_div_z"), (nent, "nent")):
        assert abs(got.item() - ref[k].item()) / abs(ref[k].item()) < 1e-5, k
    fused = m.run_model(x, x, None, eps=eps, gumbel_u=u)
    assert abs(fused.item() - loss.item()) / abs(loss.item()) < 1e-5
