"""GPU bring-up diagnostics: every case runs in its own subprocess (a trapped kernel poisons the
CUDA context) with a timeout, and one JSON line per case is appended to gpurun_out/diag.jsonl.
Test infrastructure (imports the oracle)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "diag.jsonl")


def worker(kind, arg):
    import torch
    from tests.helpers import CONFIGS, make_engine, run_parity
    if kind == "gemm":
        impl, M, N, K, ta, tb, split = arg
        eng = make_engine(CONFIGS["tiny_vae"], "bf16")
        g = torch.Generator().manual_seed(1)
        A = torch.randn((K, M) if ta else (M, K), generator=g).to(torch.bfloat16).float()
        B = torch.randn((N, K) if tb else (K, N), generator=g).to(torch.bfloat16).float()
        t0 = time.time()
        C = eng.debug_gemm(impl, A, B, bool(ta), bool(tb), split).cpu().double()
        R = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
        err = ((C - R).norm() / R.norm()).item()
        # where are the wrong entries?  (helps decode descriptor mistakes)
        bad = ((C - R).abs() > 1e-3 * R.abs().max()).nonzero()
        info = {"err": err, "nbad": int(bad.shape[0]), "sec": time.time() - t0}
        if bad.shape[0]:
            info["bad_rows"] = sorted(set(bad[:, 0].tolist()))[:12]
            info["bad_cols"] = sorted(set(bad[:, 1].tolist()))[:12]
            info["c00"] = C[:2, :4].tolist(); info["r00"] = R[:2, :4].tolist()
        return info
    if kind == "parity":
        name, precision = arg
        terr, gerr = run_parity(CONFIGS[name], precision)
        worst = max(gerr, key=gerr.get)
        return {"terms": terr, "grad_max": gerr[worst], "grad_worst": worst,
                "grads": {k: float("%.3g" % v) for k, v in gerr.items()}}
    raise ValueError(kind)


def run_case(kind, arg, env_extra=None, timeout=180):
    env = dict(os.environ)
    env.update(env_extra or {})
    t0 = time.time()
    try:
        r = subprocess.run([sys.executable, __file__, "worker", kind, json.dumps(arg)], capture_output=True, text=True,
                           timeout=timeout, env=env)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        res = json.loads(line[-1][7:]) if line else {"error": (r.stderr or r.stdout)[-600:], "rc": r.returncode}
    except subprocess.TimeoutExpired:
        res = {"error": "timeout"}
    rec = {"kind": kind, "arg": arg, "env": env_extra or {}, "wall": round(time.time() - t0, 1), **res}
    with open(OUT, "a") as f:
        f.write(json.dumps(rec) + "\n")
    brief = {k: v for k, v in rec.items() if k not in ("grads", "c00", "r00")}
    print(json.dumps(brief), flush=True)
    return rec


def main():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    stage = sys.argv[1] if len(sys.argv) > 1 else "all"
    if stage in ("all", "gemm"):
        run_case("gemm", [0, 100, 512, 784, 0, 0, 1])
        for M, N, K in [(128, 128, 64), (128, 64, 128), (100, 512, 784), (256, 784, 512), (300, 128, 512), (129, 1024, 64),
                        (16384, 512, 512), (100, 112, 72)]:
            run_case("gemm", [1, M, N, K, 0, 1, 1])
        for M, N, K, s in [(128, 128, 64, 1), (128, 64, 128, 1), (512, 512, 100, 1), (784, 512, 100, 2), (512, 784, 256, 3),
                           (64, 512, 1000, 4), (512, 128, 16384, 37), (200, 136, 333, 2)]:
            run_case("gemm", [1, M, N, K, 1, 0, s])
    if stage in ("all", "fp32"):
        for name in ["tiny_vae", "tiny_gmp", "tiny_gmvae", "nohidden_gmvae", "cfg1", "cfg2", "cfg3", "run_train_sh"]:
            run_case("parity", [name, "fp32"])
    if stage in ("all", "bf16"):
        for flags in ["1", "3", "6", "4", "2", "0"]:
            for name in ["tiny_gmvae", "cfg3"]:
                run_case("parity", [name, "bf16"], {"GMVAE_DEBUG_FLAGS": flags})
        for name in ["tiny_vae", "tiny_gmp", "nohidden_gmvae", "cfg1", "cfg2", "run_train_sh"]:
            run_case("parity", [name, "bf16"])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "worker":
        res = worker(sys.argv[2], json.loads(sys.argv[3]))
        print("RESULT " + json.dumps(res))
    else:
        main()
