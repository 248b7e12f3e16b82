"""The CLI keeps the reference's 18 flags, defaults and --flag=value syntax (run_gmvae.py:11-58);
the runner keeps create_model's fixed hyper-parameters and the logdir layout (the data pipeline, hooks, checkpoints
and the train / eval loops: tests/test_host_cpu.py)."""
import types

from gmvae_b200 import run_gmvae, runners


def test_flags_and_defaults():
    p = run_gmvae.build_parser()
    f = p.parse_args([])
    want = dict(mode="train", model="gmvae", latent_size=8, hidden_size=64, num_layers=1, mixture_components=10,
                batch_size=16, logdir="/tmp/smc_vi", random_seed=None, learning_rate=0.001, max_steps=int(1e9),
                early_stop_rounds=1000, early_stop_threshold=0.001, summarise_every=50, gpu_id="0", gpu_num="0",
                num_samples=10, num_generations=10, split="train", dataset_path=None, image_summaries=1)
    for k, v in want.items():
        assert getattr(f, k) == v, k
    f = p.parse_args(["--mode=train", "--model=vae_gmp", "--latent_size=128", "--hidden_size=512", "--batch_size=64",
                      "--logdir=/x", "--summarise_every=1000", "--early_stop_rounds=5000", "--learning_rate=0.001",
                      "--gpu_id=0", "--gpu_num=0"])                    # bin/run_train.sh:5-13
    assert f.model == "vae_gmp" and f.latent_size == 128 and f.batch_size == 64


def test_create_model_hyperparameters_and_logdir():
    cfg = types.SimpleNamespace(model="gmvae", latent_size=64, hidden_size=512, num_layers=2, mixture_components=10,
                                logdir="/tmp/l")
    m = runners.create_model(cfg, 784)
    kw = m._engine_kwargs()
    assert kw["hidden_sizes"] == [512, 512] and kw["sigma_min"] == 0.0 and kw["raw_sigma_bias"] == 0.5
    assert kw["temperature"] == 1.0 and kw["mixture_components"] == 10
    assert runners.logdir_for(cfg) == "/tmp/l/gmvae/h512_n2_z64"        # runners.py:212-217 (K is not in the path)
    cfg.model = "vae"
    assert runners.create_model(cfg, 784)._engine_kwargs()["model"] == "vae"
    cfg.model = "vae_gmp"
    assert runners.create_model(cfg, 784)._engine_kwargs()["model"] == "vae_gmp"
