"""Throughput of the reference-facing driver itself: `run_gmvae --mode=train` (gmvae_b200/runners.run_train) at the
cfg4 shape on one GPU -- device-resident synthetic MNIST-sized set, every batch binarised on the device
(gmvae_binarize), eager training steps, loss window read back every 200 steps, no image summaries.  Wall clock
around the steady part of the loop (the summaries' own `global_step/sec`).  Prints one JSON line."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gmvae_b200 import run_gmvae, runners  # noqa: E402


def main():
    batch = int(os.environ.get("BATCH", "16384"))
    logdir = tempfile.mkdtemp(prefix="gmvae_loop_")
    cfg = run_gmvae.build_parser().parse_args(
        ["--mode=train", "--model=gmvae", "--latent_size=64", "--hidden_size=512", "--num_layers=2", f"--batch_size={batch}",
         f"--logdir={logdir}", "--random_seed=3", "--summarise_every=200", "--max_steps=1999", "--image_summaries=0",
         "--early_stop_rounds=100000"])
    eng = runners.run_train(cfg)
    recs = [json.loads(l) for l in open(os.path.join(runners.logdir_for(cfg), "summaries.jsonl"))]
    rates = sorted(r["global_step/sec"] for r in recs[2:])              # skip the first windows (allocation, warm-up)
    steps_per_s = rates[len(rates) // 2]
    n = 60000
    rows_per_step = n / ((n + batch - 1) // batch)                       # the short last batch of every pass counts as a step
    print(json.dumps({"driver": "run_gmvae --mode=train (runners.run_train)", "batch_size": batch, "global_step": eng.global_step,
                      "steps_per_s_median_window": steps_per_s, "ms_per_step": 1e3 / steps_per_s,
                      "mean_rows_per_step": rows_per_step, "samples_per_s": steps_per_s * rows_per_step,
                      "elbo_first_last": [recs[0]["elbo"], recs[-1]["elbo"]],
                      "note": "synthetic 60000 x 784 set resident in HBM, binarised on the device every step; eager steps (no CUDA graph); "
                              "one device->host read per 200 steps; cluster_acc computed at every summary"}))
    eng.close()


if __name__ == "__main__":
    main()
