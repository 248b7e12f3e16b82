#!/usr/bin/env python
"""dump_reference.py -- run the UNMODIFIED reference (mazrk7/gmvae) on fixed weights, inputs and injected sampling noise and
dump everything a parity test needs to one .npz.  SURVEY.md section 4 item 6: the step that pins the oracle to the reference.

It cannot run in this repository's image (TensorFlow 1.13.1 / TFP 0.6.0 / Sonnet v1 need Python <= 3.7 and are not
installed): run it on any box that has the reference's pinned dependencies (README.md:19-23 of the reference),

    python tools/dump_reference.py --reference /path/to/gmvae/scripts --model gmvae --latent_size 64 --hidden_size 512 \
        --num_layers 2 --mixture_components 10 --batch_size 100 --out tests/golden/reference/cfg3.npz

and commit the .npz under tests/golden/reference/.  tests/test_reference_dump.py picks up every file there and compares the
oracle (CPU) and the CUDA step (GPU) with it; with no file present those tests say so and skip.

What it does (TF-1.x graph mode, nothing of this repository is imported):
  1. builds the model exactly as runners.create_model does (runners.py:65-103) and calls model.run_model on a bool
     placeholder [B, 784] (runners.py:113-134 without the dataset);
  2. assigns every trainable variable from a numpy generator (Xavier-scale weights, small non-zero biases) and stores the
     values under their TF names (`encoder_y_fcnet/linear_0/w` ...);
  3. finds the sampling ops of the loss sub-graph -- the RandomStandardNormal inside MultivariateNormalDiag.sample
     (gmvae.py:248 / vae.py:171) and the RandomUniform inside RelaxedOneHotCategorical.sample (gmvae.py:240) -- and FEEDS
     their outputs (feed_dict accepts any tensor), so the run is deterministic and the noise is known;
  4. fetches loss, the scalar summaries nll_scalar / kl_div_z / nent, and tf.gradients(loss, tf.trainable_variables())
     (what opt.compute_gradients returns, runners.py:182), plus one Adam step's updated variables (runners.py:181-183).

.npz layout: `config` (json), `x` uint8 [B,D], `eps` float32 [B,Z], `u` float32 [B,K] (GMVAE), `param/<name>`,
`term/{loss,nll,kl_div_z,nent}`, `grad/<name>`, `adam1/<name>` (variables after one AdamOptimizer(lr) step from m = v = 0).
"""
from __future__ import absolute_import, division, print_function

import argparse
import json
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="the reference's scripts/ directory")
    ap.add_argument("--model", default="gmvae", choices=["gmvae", "vae", "vae_gmp"])
    ap.add_argument("--latent_size", type=int, default=64)
    ap.add_argument("--hidden_size", type=int, default=512)
    ap.add_argument("--num_layers", type=int, default=2)
    ap.add_argument("--mixture_components", type=int, default=10)
    ap.add_argument("--batch_size", type=int, default=100)
    ap.add_argument("--learning_rate", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=2024)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()

    sys.path.insert(0, args.reference)
    try:                             # utils.cluster_acc uses the Python-2 builtin `xrange` (utils.py:179): under
        import builtins              # Python 3 provide the name instead of editing the reference
        if not hasattr(builtins, "xrange"):
            builtins.xrange = range
    except ImportError:
        pass
    import tensorflow as tf          # 1.13.1
    import runners                   # the reference's own module (scripts/runners.py)

    D, B, Z, K = 784, args.batch_size, args.latent_size, args.mixture_components
    rng = np.random.RandomState(args.seed)
    p = rng.uniform(size=D)
    x = (rng.uniform(size=(B, D)) < p)
    eps = rng.standard_normal(size=(B, Z)).astype(np.float32)
    u = np.maximum(rng.uniform(size=(B, K)), np.finfo(np.float32).tiny).astype(np.float32)

    with tf.Graph().as_default() as graph:
        images = tf.placeholder(tf.bool, [B, D], name="images")
        labels = tf.placeholder(tf.int64, [B], name="labels")
        model = runners.create_model(args, data_dim=D)                 # args carries the same attribute names as FLAGS
        n_ops_before = len(graph.get_operations())
        if args.model == "gmvae":
            loss = model.run_model(images, images, labels)
        else:
            loss = model.run_model(images, images)
        loss_ops = graph.get_operations()[n_ops_before:]
        variables = tf.trainable_variables()
        grads = tf.gradients(loss, variables)
        opt = tf.train.AdamOptimizer(args.learning_rate)
        train_op = opt.apply_gradients(list(zip(grads, variables)))

        # the sampling ops of the loss sub-graph, in graph order
        normals = [op for op in loss_ops if op.type == "RandomStandardNormal"]
        uniforms = [op for op in loss_ops if op.type == "RandomUniform"]
        assert len(normals) == 1, "expected one q_z.sample(): %r" % [o.name for o in normals]
        assert len(uniforms) == (1 if args.model == "gmvae" else 0), [o.name for o in uniforms]
        feed = {images: x, labels: np.zeros([B], np.int64)}
        feed[normals[0].outputs[0]] = eps.reshape(normals[0].outputs[0].shape.as_list())
        if uniforms:
            # random_uniform(minval=tiny, maxval=1) = RandomUniform * (1 - tiny) + tiny; feeding the raw op with u reproduces
            # u to within tiny * (1 - u) < 1.2e-38
            feed[uniforms[0].outputs[0]] = u.reshape(uniforms[0].outputs[0].shape.as_list())

        # the scalar summaries the reference attaches (gmvae.py:255-268, vae.py:178-186)
        terms = {"loss": loss}
        for op in graph.get_operations():
            if op.type == "ScalarSummary":
                tag = op.name.split("/")[-1]
                if tag in ("nll_scalar", "kl_div_z", "nent"):
                    terms["nll" if tag == "nll_scalar" else tag] = op.inputs[1]

        with tf.Session(config=tf.ConfigProto(intra_op_parallelism_threads=1, inter_op_parallelism_threads=1)) as sess:
            sess.run(tf.global_variables_initializer())
            params = {}
            for v in variables:
                name = v.name.split(":")[0]
                shape = v.shape.as_list()
                if name.endswith("/b"):
                    val = 0.05 * rng.standard_normal(size=shape)
                else:
                    fan = (shape[0], shape[0]) if len(shape) == 1 else (shape[0], shape[1])
                    lim = np.sqrt(6.0 / (fan[0] + fan[1]))
                    val = rng.uniform(-lim, lim, size=shape)
                params[name] = val.astype(np.float32)
                v.load(params[name], sess)
            names = sorted(terms)
            fetched = sess.run([terms[k] for k in names] + grads, feed_dict=feed)
            out = {"config": json.dumps({k: getattr(args, k) for k in ("model", "latent_size", "hidden_size", "num_layers",
                                                                       "mixture_components", "batch_size", "learning_rate", "seed")}),
                   "x": x.astype(np.uint8), "eps": eps}
            if uniforms:
                out["u"] = u
            for k, val in zip(names, fetched[:len(names)]):
                out["term/" + k] = np.float64(val)
            if "nent" not in terms:
                out["term/nent"] = np.float64(0.0)
            for v, g in zip(variables, fetched[len(names):]):
                out["grad/" + v.name.split(":")[0]] = np.asarray(g, np.float32)
            for name, val in params.items():
                out["param/" + name] = val
            sess.run(train_op, feed_dict=feed)                         # one TF-form Adam step from m = v = 0
            for v in variables:
                out["adam1/" + v.name.split(":")[0]] = sess.run(v)
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, {k: float(out["term/" + k]) for k in ("loss", "nll", "kl_div_z", "nent")})


if __name__ == "__main__":
    main()
