"""gmvae_b200 -- B200-native training step of mazrk7/gmvae behind the reference's Python surface.

Host layer only: PyTorch owns device memory and streams; all arithmetic of the step runs in
hand-written sm_100a CUDA kernels inside libgmvae_b200.so (include/gmvae_abi.h)."""
from ._lib import load as load_library  # noqa: F401
from .engine import Engine  # noqa: F401
from .base import ConditionalBernoulli, ConditionalCategorical, ConditionalNormal  # noqa: F401
from .vae import VAE, TrainableVAE, create_vae  # noqa: F401
from .gmvae import GMVAE, TrainableGMVAE, create_gmvae  # noqa: F401

__all__ = ["Engine", "ConditionalNormal", "ConditionalBernoulli", "ConditionalCategorical", "VAE", "TrainableVAE",
           "create_vae", "GMVAE", "TrainableGMVAE", "create_gmvae", "load_library"]
