"""Host-side mirrors of the reference's data pipeline, hooks, checkpoints, summaries and the train / eval
drivers: gmvae_b200/{data,utils,runners}.py.  No GPU: the engine and the model are
replaced by fakes, so what is tested is the host logic -- ordering, sharding, the early-stopping state machine,
file layouts, the eval loop's averaging -- not arithmetic."""
import gzip
import json
import os
import types

import numpy as np
import pytest
import torch

from gmvae_b200 import data, runners, utils


# ---------------------------------------------------------------------------- data.py
def test_idx_roundtrip_and_errors(tmp_path):
    img = np.random.default_rng(0).integers(0, 256, size=(7, 28, 28), dtype=np.uint8)
    for name in ("a-idx3-ubyte", "a-idx3-ubyte.gz"):
        p = str(tmp_path / name)
        data.write_idx(p, img)
        got = data.read_idx(p)
        assert got.dtype == np.uint8 and got.shape == (7, 28, 28) and (got == img).all()
    bad = tmp_path / "bad"
    bad.write_bytes(b"\x01\x00\x08\x01\x00\x00\x00\x01\x00")
    with pytest.raises(ValueError, match="not an IDX"):
        data.read_idx(str(bad))
    trunc = tmp_path / "trunc"
    trunc.write_bytes(bytes([0, 0, 8, 1]) + (5).to_bytes(4, "big") + b"\x00\x01")
    with pytest.raises(ValueError, match="payload"):
        data.read_idx(str(trunc))
    with gzip.open(str(tmp_path / "hdr.gz"), "wb") as f:
        f.write(bytes([0, 0, 8, 3]) + b"\x00")
    with pytest.raises(ValueError, match="truncated"):
        data.read_idx(str(tmp_path / "hdr.gz"))


def test_load_mnist_from_idx_dir_and_npz(tmp_path):
    rng = np.random.default_rng(1)
    xtr, ytr = rng.integers(0, 256, size=(20, 28, 28), dtype=np.uint8), rng.integers(0, 10, size=20).astype(np.uint8)
    xte, yte = rng.integers(0, 256, size=(6, 28, 28), dtype=np.uint8), rng.integers(0, 10, size=6).astype(np.uint8)
    d = tmp_path / "idx"
    d.mkdir()
    data.write_idx(str(d / "train-images-idx3-ubyte.gz"), xtr)
    data.write_idx(str(d / "train-labels-idx1-ubyte.gz"), ytr)
    data.write_idx(str(d / "t10k-images.idx3-ubyte"), xte)          # the other common spelling
    data.write_idx(str(d / "t10k-labels.idx1-ubyte"), yte)
    x, y = data.load_mnist(str(d), "train")
    assert x.dtype == np.uint8 and x.shape == (20, 784) and y.dtype == np.int64 and (y == ytr).all()
    assert (x == xtr.reshape(20, 784)).all()
    x, y = data.load_mnist(str(d), "test")
    assert x.shape == (6, 784) and (y == yte).all()
    n = tmp_path / "npz"
    n.mkdir()
    np.savez(str(n / "mnist.npz"), x_train=xtr, y_train=ytr, x_test=xte, y_test=yte)
    x, y = data.load_mnist(str(n), "test")
    assert (x == xte.reshape(6, 784)).all() and (y == yte).all()
    x, _ = data.load_mnist(str(n / "mnist.npz"), "train")
    assert x.shape == (20, 784)
    assert data.load_mnist(None, "train") is None and data.load_mnist(str(tmp_path), "train") is None
    with pytest.raises(ValueError):
        data.load_mnist(str(d), "validation")


def test_synthetic_mnist_contract():
    x, y = data.synthetic_mnist("test", num_examples=500)
    assert x.dtype == np.uint8 and x.shape == (500, 784) and y.dtype == np.int64 and y.shape == (500,)
    assert set(np.unique(y)) <= set(range(10)) and x.max() > 100 and np.median(x) < 64
    x2, y2 = data.synthetic_mnist("test", num_examples=500)
    assert (x == x2).all() and (y == y2).all()
    assert data.synthetic_mnist("train", num_examples=8)[0].shape == (8, 784)


def test_batch_schedule_is_batch_level_shuffle():
    s = data.BatchSchedule(1050, 100, shuffle=True, repeat=False, seed=3)
    b = list(s)
    assert len(b) == 11 and sorted(b) == [(i * 100, 100) for i in range(10)] + [(1000, 50)]
    assert b != sorted(b)                                            # shuffled ...
    assert all(first % 100 == 0 for first, _ in b)                   # ... but whole batches only (runners.py:50-57)
    assert list(data.BatchSchedule(250, 100, shuffle=False, repeat=False)) == [(0, 100), (100, 100), (200, 50)]
    it = iter(data.BatchSchedule(250, 100, shuffle=False, repeat=True))
    assert [next(it) for _ in range(7)] == [(0, 100), (100, 100), (200, 50)] * 2 + [(0, 100)]
    with pytest.raises(ValueError):
        data.BatchSchedule(0, 10, False, False)


class _Binarizer:
    def __init__(self):
        self.calls = []

    def __call__(self, intensities, batch=None, first_row=0, row_index=None, draw=0, out=None):
        self.calls.append((first_row, batch, draw, out is not None))
        x = (intensities[first_row:first_row + batch] < 128).to(torch.uint8)
        if out is not None:
            out.copy_(x)
            return out
        return x


def test_device_dataset_draws_and_static_buffer():
    inten = torch.arange(250 * 4, dtype=torch.int64).reshape(250, 4).remainder(256).to(torch.uint8)
    labels = torch.arange(250)
    bz = _Binarizer()
    static = torch.zeros(100, 4, dtype=torch.uint8)
    ds = data.DeviceDataset(inten, labels, 100, shuffle=False, repeat=False, binarize=bz, first_draw=40, static_out=static)
    got = list(ds)
    assert [tuple(x.shape) for x, _ in got] == [(100, 4), (100, 4), (50, 4)]
    assert [c[2] for c in bz.calls] == [40, 41, 42]                  # a new draw counter per batch
    assert [c[3] for c in bz.calls] == [True, True, False]           # full batches land in the static buffer
    assert got[0][0].data_ptr() == static.data_ptr() and (got[2][1] == labels[200:250]).all()
    assert ds.last_global_rows == 50 and ds.num_examples == 250
    with pytest.raises(ValueError):
        data.DeviceDataset(inten, labels[:10], 100, False, False, bz)


def test_device_dataset_shards_global_batches_by_rank():
    inten = torch.zeros(205, 4, dtype=torch.uint8)
    labels = torch.arange(205)
    seen = {}
    for rank in range(4):
        bz = _Binarizer()
        ds = data.DeviceDataset(inten, labels, 100, shuffle=True, repeat=False, binarize=bz, seed=9, world=4, rank=rank)
        seen[rank] = [(lab.tolist(), ds.last_global_rows) for _, lab in ds]
    n_batches = len(seen[0])
    assert n_batches == 3 and all(len(v) == 3 for v in seen.values())
    for i in range(n_batches):
        rows = sum((seen[r][i][0] for r in range(4)), [])
        g = seen[0][i][1]
        assert all(seen[r][i][1] == g for r in range(4)) and len(rows) == g
        assert rows == list(range(rows[0], rows[0] + g))              # contiguous shards, rank order, nothing lost
    # a global batch with fewer rows than ranks is skipped by every rank (nobody may sit out an all-reduce)
    ds = data.DeviceDataset(inten[:102], labels[:102], 100, False, False, _Binarizer(), seed=0, world=4, rank=3)
    assert [lab.tolist() for _, lab in ds] == [list(range(75, 100))]
    with pytest.raises(ValueError, match="seed"):
        data.DeviceDataset(inten, labels, 100, True, True, _Binarizer(), world=2, rank=0)


# ---------------------------------------------------------------------------- utils.py
def test_early_stopping_hook_state_machine():
    """utils.py:13-57: the first call only initialises; the counter counts calls since the last improvement."""
    h = utils.EarlyStoppingHook(max_steps=3, threshold=0.1)
    assert not h.after_run(100.0, 1) and h._prev_loss is None        # first call: reset only (:38-45)
    assert not h.after_run(100.0, 2) and h._prev_loss == 100.0 and h._steps == 0
    assert not h.after_run(95.0, 3)                                  # not a 10 % improvement: 1
    assert not h.after_run(89.0, 4) and h._steps == 0                # improvement: reset
    assert [h.after_run(88.0, s) for s in (5, 6)] == [False, False]
    assert h.after_run(88.0, 7) and h.stop_step == 7                 # third stale step
    # the global step going backwards (recovery) resets everything
    h = utils.EarlyStoppingHook(max_steps=2, threshold=0.0)
    for s, l in ((10, 5.0), (11, 5.0), (12, 5.0)):
        h.after_run(l, s)
    assert h._steps == 1
    h.after_run(5.0, 3)
    assert h._steps == 0 and h._prev_loss is None and h._last_step == 3 and not h.stop_requested
    assert utils.summary_formatter({"step": 50, "loss": 1.5}) == "Step 50, loss: 1.500000"


class _FakeEngineState:
    def __init__(self):
        self.loaded = None

    def load_state_dict(self, sd):
        self.loaded = sd


def test_checkpoint_layout_and_restore(tmp_path):
    logdir = str(tmp_path / "run")
    e = _FakeEngineState()
    assert utils.get_checkpoint_state(logdir) is None and not utils.restore_checkpoint_if_exists(e, logdir)
    for step in range(0, 70, 10):
        utils.save_checkpoint({"w": torch.full((2,), float(step)), "global_step": torch.tensor(step)}, logdir, step)
    st = utils.get_checkpoint_state(logdir)
    assert st["model_checkpoint_path"] == "model.ckpt-60"
    assert st["all_model_checkpoint_paths"] == [f"model.ckpt-{s}" for s in (20, 30, 40, 50, 60)]      # max_to_keep = 5
    files = sorted(f for f in os.listdir(logdir) if f.startswith("model.ckpt"))
    assert files == sorted(f"model.ckpt-{s}" for s in (20, 30, 40, 50, 60))
    assert open(os.path.join(logdir, "checkpoint")).readline() == 'model_checkpoint_path: "model.ckpt-60"\n'
    assert utils.restore_checkpoint_if_exists(e, logdir) and int(e.loaded["global_step"]) == 60
    with pytest.raises(TimeoutError):
        utils.wait_for_checkpoint(e, str(tmp_path / "empty"), poll_secs=0.01, max_wait=0.03)
    utils.wait_for_checkpoint(e, logdir, poll_secs=0.01, max_wait=0.0)


def test_pack_images_matches_explicit_tiling():
    imgs = np.arange(7 * 3 * 2 * 1, dtype=np.float32).reshape(7, 3, 2, 1)
    tile = utils.pack_images(imgs, rows=2, cols=3)
    assert tile.shape == (1, 2 * 3, 3 * 2, 1)
    for r in range(2):
        for c in range(3):
            assert (tile[0, r * 3:(r + 1) * 3, c * 2:(c + 1) * 2] == imgs[r * 3 + c]).all()
    # rows = min(rows, batch); cols = min(batch // rows, cols)   (utils.py:124-125)
    assert utils.pack_images(imgs, rows=8, cols=8).shape == (1, 7 * 3, 1 * 2, 1)
    assert utils.pack_images(imgs, rows=5, cols=5).shape == (1, 5 * 3, 1 * 2, 1)
    x = torch.zeros(4, 784)
    assert tuple(utils.unflatten_tensor(x, (28, 28, 1)).shape) == (4, 28, 28, 1)
    assert tuple(utils.flatten_tensor(torch.zeros(4, 28, 28, 1), (28, 28, 1)).shape) == (4, 784)


def test_mode_entropy_cluster_acc():
    assert utils.mode_tensor(np.array([3, 1, 1, 3, 2])) == 3.0       # tie -> first occurrence (unique_with_counts order)
    assert utils.mode_tensor(np.array([1, 3, 3, 1, 1])) == 1.0
    logits = torch.tensor([[1.0, 2.0, 0.5], [0.0, 0.0, 0.0]])
    p = torch.softmax(logits, 1)
    want = -(p * torch.log(p)).sum(1)
    assert torch.allclose(utils.entropy(logits, p), want, atol=1e-6)
    assert abs(float(utils.entropy(logits, p)[1]) - np.log(3)) < 1e-6
    # clusters 0/1/2 hold labels {7,7,4}, {4,4}, {} -> mapped to 7, 4, (0): 4 of 5 right
    onehot = np.eye(3)[[0, 0, 0, 1, 1]] * 5.0
    labels = np.array([7, 7, 4, 4, 4])
    assert utils.cluster_acc(onehot, labels, 3) == pytest.approx(4 / 5)
    # an empty cluster contributes mode 0.0 (utils.py:183-185); label 0 in a non-empty cluster still counts
    assert utils.cluster_acc(np.eye(2)[[0, 0]], np.array([0, 0]), 2) == 1.0
    assert utils.cluster_acc(torch.tensor(onehot), torch.tensor(labels), 3) == pytest.approx(4 / 5)


def test_plots_and_tiles_are_pngs(tmp_path):
    from PIL import Image
    w = utils.SummaryWriter(str(tmp_path / "s"))
    p = utils.image_tile_summary(w, "inputs", torch.rand(30, 28, 28, 1), step=50, rows=5, cols=5)
    assert p.endswith("image_summaries/inputs/step_50.png") and Image.open(p).size == (140, 140)
    w.scalars(50, {"elbo": -1.5})
    w.close()
    assert json.loads(open(str(tmp_path / "s" / "summaries.jsonl")).readline()) == {"step": 50, "elbo": -1.5}
    utils.scatter_png(str(tmp_path / "lat"), np.random.default_rng(0).normal(size=(200, 2)), np.arange(200) % 10, size=200)
    im = Image.open(str(tmp_path / "lat.png"))
    assert im.size == (200, 200) and len(im.getcolors(100000)) > 5
    utils.scatter_png(str(tmp_path / "one"), np.zeros((1, 2)), size=64)         # degenerate extent
    assert utils.display_images(str(tmp_path / "grid"), np.random.rand(100, 28, 28, 1)) == 10
    assert Image.open(str(tmp_path / "grid.png")).size == (280, 280)
    assert utils.display_images(str(tmp_path / "grid3"), np.random.rand(10, 28, 28, 1)) == 3     # reference: IndexError
    z = utils.reduce_dimensionality(np.random.default_rng(0).normal(size=(40, 5)), random_state=0)
    assert z.shape == (40, 2)
    two = np.random.rand(9, 2)
    assert utils.reduce_dimensionality(two) is two or (utils.reduce_dimensionality(two) == two).all()


# ---------------------------------------------------------------------------- runners.py with a fake engine
class FakeEngine:
    def __init__(self, loss_fn, data_size=784):
        self.device, self.data_size, self._step, self.loss_fn = torch.device("cpu"), data_size, 0, loss_fn
        self.draws, self.global_batches = [], []

    def binarize(self, intensities, batch=None, first_row=0, row_index=None, draw=0, out=None):
        self.draws.append(draw)
        return (intensities[first_row:first_row + batch] < 128).to(torch.uint8)

    def train_step(self, x, eps=None, gumbel_u=None, global_batch=None):
        assert x.dtype == torch.uint8 and x.dim() == 2 and x.shape[1] == self.data_size
        self._step += 1
        self.global_batches.append(global_batch)
        l = float(self.loss_fn(self._step))
        return torch.tensor([l, l + 2.0, -1.0, -1.0])

    @property
    def global_step(self):
        return self._step

    def state_dict(self):
        return {"w": torch.zeros(2), "global_step": torch.tensor(self._step)}

    def load_state_dict(self, sd):
        self._step = int(sd["global_step"])


class FakeModel:
    def __init__(self, K=10, Z=4, eval_loss=lambda n: 100.0):
        self.K, self.Z, self.eval_loss, self.random_seed = K, Z, eval_loss, None
        self.calls = []

    def encoder_y_logits(self, x):
        return torch.eye(self.K)[torch.arange(x.shape[0]) % self.K]

    def reconstruct_images(self, x):
        return x.float()

    def generate_samples(self, num_samples, clusters=None):
        n = num_samples * (self.K if clusters is None else len(clusters))
        self.calls.append(("generate_samples", num_samples, clusters))
        return torch.randn(n, self.Z)

    def generate_sample_images(self, z=None, num_samples=1, name="sample_images"):
        if z is None:
            z = self.generate_samples(num_samples)
        return torch.rand(z.shape[0], 784)

    def transform(self, x):
        return torch.randn(x.shape[0], self.Z)

    def run_model(self, images, targets, labels=None):
        return torch.tensor(self.eval_loss(images.shape[0]))


def _config(tmp_path, **kw):
    base = dict(mode="train", model="gmvae", latent_size=4, hidden_size=8, num_layers=1, mixture_components=10, batch_size=100,
                logdir=str(tmp_path / "logs"), random_seed=None, learning_rate=1e-3, max_steps=10 ** 9, early_stop_rounds=1000,
                early_stop_threshold=0.001, summarise_every=50, gpu_id="0", gpu_num="0", num_samples=10, num_generations=10,
                split="test", dataset_path=None, image_summaries=1)
    base.update(kw)
    return types.SimpleNamespace(**base)


def _patch(monkeypatch, engine, model):
    monkeypatch.setattr(runners, "create_model", lambda config, data_dim: model)
    monkeypatch.setattr(runners, "_configure", lambda m, config, device: engine)
    monkeypatch.setattr(data, "SPLIT_SIZES", {"train": 1000, "test": 250})      # small synthetic stand-in
    monkeypatch.delenv("GMVAE_MNIST_DIR", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)


def test_create_dataset_contract_and_errors(tmp_path, monkeypatch):
    eng = FakeEngine(lambda s: 1.0)
    _patch(monkeypatch, eng, FakeModel())
    cfg = _config(tmp_path, batch_size=32)
    it = runners.create_dataset(cfg, "train", shuffle=True, repeat=True, engine=eng)
    img, lab = next(it)
    assert img.dtype == torch.uint8 and tuple(img.shape) == (32, 28, 28, 1) and int(img.max()) <= 1     # runners.py:32-35
    assert lab.dtype == torch.int64 and tuple(lab.shape) == (32,) and it.last_global_rows == 32 and it.num_examples == 1000
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        runners.create_dataset(cfg, "train", True, True)
    cfg.dataset_path = str(tmp_path / "nothing_here")
    os.makedirs(cfg.dataset_path)
    with pytest.raises(FileNotFoundError):
        runners.create_dataset(cfg, "train", True, True, engine=eng)


def test_run_train_loop_summaries_early_stop_and_resume(tmp_path, monkeypatch, capsys):
    # loss improves for 120 steps, then is flat: with early_stop_rounds=100 the hook fires in the window ending at 250
    eng = FakeEngine(lambda s: 1000.0 - 5.0 * min(s, 120))
    model = FakeModel()
    _patch(monkeypatch, eng, model)
    cfg = _config(tmp_path, early_stop_rounds=100, summarise_every=50)
    runners.run_train(cfg)
    out = capsys.readouterr().out
    assert "Step 50, loss: 750.000000" in out and "[Early Stopping Criterion Satisfied]" in out
    # flat from step 121 on: steps 121..220 are 100 calls without improvement -> requested at step 220,
    # noticed at the end of that window (step 250)
    assert eng.global_step == 250
    logdir = runners.logdir_for(cfg)
    assert logdir.endswith("logs/gmvae/h8_n1_z4")
    recs = [json.loads(l) for l in open(os.path.join(logdir, "summaries.jsonl"))]
    assert [r["step"] for r in recs] == [50, 100, 150, 200, 250]
    assert set(recs[0]) == {"step", "elbo", "nll_scalar", "kl_div_z", "nent", "cluster_acc", "global_step/sec"}
    assert recs[0]["elbo"] == -750.0 and recs[0]["nll_scalar"] == 752.0 and 0.0 <= recs[0]["cluster_acc"] <= 1.0
    for name in ("inputs", "reconstructions", "samples"):
        assert os.path.exists(os.path.join(logdir, "image_summaries", name, "step_250.png"))
    assert ("generate_samples", 1, None) in model.calls                # GMVAE: one sample per component (runners.py:133)
    assert utils.get_checkpoint_state(logdir)["model_checkpoint_path"] == "model.ckpt-250"
    assert eng.draws == list(range(250)) and set(eng.global_batches) == {100}
    # a second run restores step 250 and runs to max_steps (loop condition `cur_step <= max_steps`, runners.py:231)
    eng2 = FakeEngine(lambda s: 1000.0 / s)
    _patch(monkeypatch, eng2, FakeModel())
    cfg2 = _config(tmp_path, max_steps=299, summarise_every=20, image_summaries=0, model="gmvae")
    runners.run_train(cfg2)
    assert eng2.global_step == 300 and eng2.draws[0] == 250
    assert "Restored checkpoint of step 250" in capsys.readouterr().out
    recs = [json.loads(l) for l in open(os.path.join(logdir, "summaries.jsonl"))][5:]
    assert [r["step"] for r in recs] == [270, 290, 300]                  # the last, short window is flushed too


def test_run_train_vae_has_no_gmvae_summaries(tmp_path, monkeypatch):
    eng, model = FakeEngine(lambda s: 10.0), FakeModel()
    _patch(monkeypatch, eng, model)
    cfg = _config(tmp_path, model="vae", max_steps=9, summarise_every=5)
    runners.run_train(cfg)
    recs = [json.loads(l) for l in open(os.path.join(runners.logdir_for(cfg), "summaries.jsonl"))]
    assert [r["step"] for r in recs] == [5, 10] and "nent" not in recs[0] and "cluster_acc" not in recs[0]
    assert ("generate_samples", 10, None) in model.calls               # VAE: ten prior samples (runners.py:136)


def test_run_eval_loop_and_outputs(tmp_path, monkeypatch):
    eng, model = FakeEngine(lambda s: 1.0), FakeModel(eval_loss=lambda n: 100.0 + n)
    _patch(monkeypatch, eng, model)
    cfg = _config(tmp_path, split="test", batch_size=100, random_seed=5)
    with pytest.raises(SystemExit):                                    # runners.py:421-423: no logdir -> exit(1)
        runners.run_eval(cfg)
    utils.save_checkpoint({"global_step": torch.tensor(1234)}, runners.logdir_for(cfg), 1234)
    res = runners.run_eval(cfg, max_wait=0.0)
    # 250 examples in batches of 100, 100, 50 -> per-batch mean losses 200, 200, 150
    assert res["step"] == 1234 and res["z"].shape == (250, 4) and res["labels"].shape == (250, 1)
    assert res["avg_loss"] == pytest.approx((200 + 200 + 150) / 250)                 # the reference's number (F10)
    assert res["loss_per_example"] == pytest.approx((200 * 100 + 200 * 100 + 150 * 50) / 250)
    sd = res["summary_dir"]
    assert sd.endswith("h8_n1_z4/test")
    for f in ("step_1234.png", "step_1234_samples.png", "step_1234_sample_images.png", "step_1234_sample_k_images.png",
              "summaries.jsonl"):
        assert os.path.exists(os.path.join(sd, f)), f
    rec = json.loads(open(os.path.join(sd, "summaries.jsonl")).readline())
    assert rec["test/loss_per_example"] == pytest.approx(2.2) and rec["step"] == 1234
    ks = [c for c in model.calls if c[2] is not None]
    assert len(ks) == 1 and ks[0][1] == 100 and len(ks[0][2]) == 1 and 0 <= ks[0][2][0] < 10    # runners.py:283-286
    assert eng.draws == [0, 1, 2]                                      # one pass, no shuffle, no repeat


# ---------------------------------------------------------------------------- run_train under torchrun (world size 2, gloo)
def _dp_worker(rank, world, port, logroot):
    import torch.distributed as dist
    from gmvae_b200 import dist as dist_mod
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    os.environ.pop("GMVAE_MNIST_DIR", None)
    torch.set_num_threads(1)

    class DPEngine(FakeEngine):
        """Like the native step under data parallelism: the loss every rank reads is the all-reduced one."""
        rows_seen = []

        def init_data_parallel(self):
            self.world_size, self.rank = dist.get_world_size(), dist.get_rank()

        def binarize(self, intensities, batch=None, first_row=0, row_index=None, draw=0, out=None):
            self.rows_seen.append((first_row, batch))
            return super().binarize(intensities, batch=batch, first_row=first_row, draw=draw)

        def train_step(self, x, eps=None, gumbel_u=None, global_batch=None):
            t = super().train_step(x, global_batch=global_batch)
            t = t + float(self.rank)                                   # a rank-dependent local term ...
            dist.all_reduce(t)                                         # ... summed over ranks, as the gradient buffer's tail is
            return t / world

    eng = DPEngine(lambda s: 500.0 - min(s, 30))
    runners.create_model = lambda config, data_dim: FakeModel()
    runners._configure = lambda m, config, device: eng
    dist_mod.init_process_group = lambda backend="nccl": dist.init_process_group("gloo", rank=rank, world_size=world)
    data.SPLIT_SIZES = {"train": 1000, "test": 100}
    cfg = types.SimpleNamespace(mode="train", model="gmvae", latent_size=4, hidden_size=8, num_layers=1, mixture_components=10,
                                batch_size=50, logdir=logroot, random_seed=4, learning_rate=1e-3, max_steps=10 ** 6,
                                early_stop_rounds=40, early_stop_threshold=0.001, summarise_every=25, gpu_id="0", gpu_num="0",
                                dataset_path=None, image_summaries=0)
    runners.run_train(cfg)
    torch.save({"step": eng.global_step, "rows": eng.rows_seen, "gb": sorted(set(eng.global_batches))}, os.path.join(logroot, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_run_train_world_size_2_gloo(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    logroot = str(tmp_path / "dp")
    os.makedirs(logroot)
    mp.spawn(_dp_worker, args=(2, port, logroot), nprocs=2, join=True)
    r0, r1 = torch.load(os.path.join(logroot, "rank0.pt")), torch.load(os.path.join(logroot, "rank1.pt"))
    # both ranks stopped at the same step: the loss they replay through the hook is the all-reduced one.
    # flat from step 31: steps 31..70 are 40 stale calls -> requested at 70, noticed at the end of that window (75)
    assert r0["step"] == r1["step"] == 75
    assert r0["gb"] == r1["gb"] == [100]                               # divisor of the batch means = world * batch_size
    # every global batch of 100 rows is split into two contiguous halves, rank order
    assert len(r0["rows"]) == len(r1["rows"]) == 75
    for (f0, n0), (f1, n1) in zip(r0["rows"], r1["rows"]):
        assert n0 == n1 == 50 and f1 == f0 + 50 and f0 % 100 == 0
    logdir = os.path.join(logroot, "gmvae", "h8_n1_z4")
    recs = [json.loads(l) for l in open(os.path.join(logdir, "summaries.jsonl"))]
    assert [r["step"] for r in recs] == [25, 50, 75]                   # written once (rank 0 only)
    assert utils.get_checkpoint_state(logdir)["all_model_checkpoint_paths"] == ["model.ckpt-75"]
