"""run_gmvae --mode=train / --mode=eval end to end on the GPU with the real engine:
device-side binarisation -> training step -> summaries, checkpoints, early-stopping replay; then the eval pass
over the test split with the forward-only helpers."""
import json
import os

import numpy as np
import pytest
import torch

from gmvae_b200 import data, run_gmvae, runners, utils

pytestmark = pytest.mark.gpu


def _flags(tmp_path, *extra):
    return run_gmvae.build_parser().parse_args(
        ["--latent_size=8", "--hidden_size=64", "--num_layers=1", "--batch_size=64", f"--logdir={tmp_path}/logs",
         "--random_seed=7", "--summarise_every=20"] + list(extra))


@pytest.mark.parametrize("model", ["gmvae", "vae_gmp", "vae"])
def test_train_then_eval(tmp_path, monkeypatch, model):
    monkeypatch.setattr(data, "SPLIT_SIZES", {"train": 2048, "test": 300})       # synthetic stand-in, small
    monkeypatch.delenv("GMVAE_MNIST_DIR", raising=False)
    cfg = _flags(tmp_path, "--mode=train", f"--model={model}", "--max_steps=59")
    eng = runners.run_train(cfg)
    assert eng.global_step == 60
    logdir = runners.logdir_for(cfg)
    recs = [json.loads(l) for l in open(os.path.join(logdir, "summaries.jsonl"))]
    assert [r["step"] for r in recs] == [20, 40, 60]
    assert all(np.isfinite(list(r.values())).all() for r in recs)
    assert recs[-1]["elbo"] > recs[0]["elbo"]                                     # 60 Adam steps do improve the bound
    assert ("cluster_acc" in recs[0]) == (model == "gmvae")
    for name in ("inputs", "reconstructions", "samples"):
        assert os.path.exists(os.path.join(logdir, "image_summaries", name, "step_60.png"))
    assert utils.get_checkpoint_state(logdir)["model_checkpoint_path"] == "model.ckpt-60"
    trained = {k: v.clone() for k, v in eng.state_dict().items()}
    eng.close()

    ecfg = _flags(tmp_path, "--mode=eval", f"--model={model}", "--split=test")
    res = runners.run_eval(ecfg, max_wait=0.0)
    assert res["step"] == 60 and res["z"].shape == (300, 8) and res["labels"].shape == (300, 1)
    assert np.isfinite(res["z"]).all() and np.isfinite(res["loss_per_example"])
    # the eval loss is in the range of the training loss, and the reference's own number is ~1/batch of it (F10)
    assert 0.4 * -recs[-1]["elbo"] < res["loss_per_example"] < 2.5 * -recs[-1]["elbo"]
    assert res["avg_loss"] < res["loss_per_example"] / 30
    for f in ("step_60.png", "step_60_samples.png", "step_60_sample_images.png"):
        assert os.path.exists(os.path.join(res["summary_dir"], f)), f
    assert os.path.exists(os.path.join(res["summary_dir"], "step_60_sample_k_images.png")) == (model == "gmvae")
    # the checkpoint restored by eval is the one training wrote
    ck = torch.load(os.path.join(logdir, "model.ckpt-60"))
    assert set(ck) == set(trained) and all(torch.equal(ck[k], trained[k]) for k in ck)


def test_train_resumes_and_stops_early(tmp_path, monkeypatch, capsys):
    monkeypatch.setattr(data, "SPLIT_SIZES", {"train": 1024, "test": 128})
    monkeypatch.delenv("GMVAE_MNIST_DIR", raising=False)
    cfg = _flags(tmp_path, "--mode=train", "--model=gmvae", "--max_steps=19", "--image_summaries=0")
    runners.run_train(cfg).close()
    cfg = _flags(tmp_path, "--mode=train", "--model=gmvae", "--max_steps=10000", "--image_summaries=0",
                 "--early_stop_rounds=30", "--early_stop_threshold=0.5")          # a 50 % improvement never happens again
    eng = runners.run_train(cfg)
    out = capsys.readouterr().out
    assert "Restored checkpoint of step 20" in out and "[Early Stopping Criterion Satisfied]" in out
    # resumed at 20; the hook's first call (step 21) only initialises, step 22 sets the reference loss, steps 23..52
    # are 30 stale steps -> requested at 52, noticed at the end of that window (step 60)
    assert eng.global_step == 60
    eng.close()
