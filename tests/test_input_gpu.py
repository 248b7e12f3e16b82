"""Input pipeline on the GPU: gmvae_binarize through the C ABI against the numpy
oracle (oracle/input_oracle.py), bit-exact; runners.py:44-47."""
import numpy as np
import pytest
import torch

from oracle import input_oracle as O

pytestmark = pytest.mark.gpu

SEED = 0x243F6A8885A308D3


def _engine(data_size=784, seed=SEED):
    import gmvae_b200
    return gmvae_b200.Engine("vae", data_size=data_size, latent_size=10, hidden_sizes=[48], max_batch=64, seed=seed)


def test_binarize_matches_oracle_contiguous_and_gathered():
    eng = _engine()
    rng = np.random.default_rng(5)
    inten = rng.integers(0, 256, size=(300, 784), dtype=np.uint8)
    dev = torch.from_numpy(inten).cuda()
    x = eng.binarize(dev, batch=100, first_row=0, draw=17)
    assert x.dtype == torch.uint8 and tuple(x.shape) == (100, 784)
    assert (x.cpu().numpy() == O.binarize(inten, None, 100, SEED, 17)).all()
    # a contiguous batch further down the dataset (the reference's shuffled batches are contiguous runs)
    x = eng.binarize(dev, batch=50, first_row=200, draw=18)
    assert (x.cpu().numpy() == O.binarize(inten[200:], None, 50, SEED, 18)).all()
    # gathered rows, repeats allowed
    idx = rng.integers(0, 300, size=77)
    x = eng.binarize(dev, row_index=torch.from_numpy(idx).cuda(), draw=19)
    assert (x.cpu().numpy() == O.binarize(inten, idx, 77, SEED, 19)).all()
    # a source that is not 4-byte aligned takes the byte path, same answer
    flat = torch.zeros(300 * 784 + 1, dtype=torch.uint8, device="cuda")
    odd = flat[1:].view(300, 784)
    odd.copy_(dev)
    assert odd.data_ptr() % 4 != 0 and odd.is_contiguous()
    x = eng.binarize(odd, batch=100, first_row=0, draw=17)
    assert (x.cpu().numpy() == O.binarize(inten, None, 100, SEED, 17)).all()
    x = eng.binarize(odd, row_index=torch.from_numpy(idx).cuda(), draw=19)
    assert (x.cpu().numpy() == O.binarize(inten, idx, 77, SEED, 19)).all()
    eng.close()


def test_binarize_edges_and_errors():
    eng = _engine(data_size=120)                              # (widths with D % 4 != 0: tests/test_input_cpu.py, same source)
    rng = np.random.default_rng(6)
    inten = rng.integers(0, 256, size=(41, 120), dtype=np.uint8)
    inten[0] = 0; inten[1] = 255
    dev = torch.from_numpy(inten).cuda()
    x = eng.binarize(dev, draw=1).cpu().numpy()
    assert (x == O.binarize(inten, None, 41, SEED, 1)).all()
    assert (x[0] == 1).all() and (x[1] == 0).all()            # inverted binarisation: 0 -> always 1, 255 -> always 0
    assert tuple(eng.binarize(dev, batch=0).shape) == (0, 120)  # empty batch: no launch, no error
    with pytest.raises(ValueError):
        eng.binarize(dev, batch=42)
    with pytest.raises(ValueError):
        eng.binarize(dev.to(torch.int32))
    with pytest.raises(RuntimeError, match="overlap"):
        eng.binarize(dev, batch=4, out=dev[4:8])
    eng.close()


def test_binarize_full_dataset_properties_and_seed():
    """MNIST-train-sized input (60 000 x 784): determinism, seed / draw sensitivity, and the reference's
    distribution P(x = 1) = 1 - intensity / 255 per intensity value."""
    eng = _engine(seed=123)
    g = torch.Generator(device="cuda").manual_seed(1)
    dev = torch.randint(0, 256, (60000, 784), dtype=torch.uint8, device="cuda", generator=g)
    a = eng.binarize(dev, draw=5)
    b = eng.binarize(dev, draw=5)
    assert torch.equal(a, b) and int(a.max()) == 1
    c = eng.binarize(dev, draw=6)
    assert float((a != c).float().mean()) > 0.2
    # first rows against the oracle at full size (same element counters regardless of the batch size)
    head = O.binarize(dev[:64].cpu().numpy(), None, 64, 123, 5)
    assert (a[:64].cpu().numpy() == head).all()
    ones = torch.zeros(256, dtype=torch.float64, device="cuda").index_add_(0, dev.reshape(-1).long(), a.reshape(-1).double())
    cnt = torch.bincount(dev.reshape(-1).long(), minlength=256).double()
    p = (ones / cnt).cpu().numpy()
    want = 1.0 - np.arange(256) / 255.0
    assert np.abs(p - want).max() < 6e-3, np.abs(p - want).max()     # ~184k draws per value: 5 sigma ~ 5.8e-3
    assert p[0] == 1.0 and p[255] == 0.0
    eng.close()


def test_binarized_batch_feeds_the_step():
    import gmvae_b200
    eng = gmvae_b200.Engine("gmvae", latent_size=8, hidden_sizes=[32], mixture_components=10, max_batch=128, seed=3)
    dev = torch.randint(0, 256, (512, 784), dtype=torch.uint8, device="cuda")
    x = eng.binarize(dev, batch=128, first_row=256, draw=0)
    loss = eng.train_step(x).cpu()
    assert torch.isfinite(loss).all() and loss[0] > 0
    eng.close()


def test_unpack_bits_matches_numpy():
    """gmvae_unpack_bits against numpy.unpackbits, D = 784 (aligned 8-byte stores) and a ragged D (tail bits, byte path)."""
    import numpy as np
    import gmvae_b200
    for D, B in ((784, 257), (203, 33)):
        eng = gmvae_b200.Engine("vae", data_size=D, latent_size=8, hidden_sizes=[16], mixture_components=1, max_batch=B, seed=1)
        x = (np.random.default_rng(D).random((B, D)) < 0.4).astype(np.uint8)
        packed = torch.from_numpy(np.packbits(x, axis=1)).cuda()
        out = eng.unpack_bits(packed)
        assert (out.cpu().numpy() == x).all()
        eng.close()
