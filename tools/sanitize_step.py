"""One small training step of every model in both precisions, for compute-sanitizer (memcheck / synccheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_step.py
The chained kernel runs with the one-launch plan forced (GMVAE_DEBUG_FLAGS=8192: row jobs, CTA pairs) as well as the default."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import gmvae_oracle as O
from tests.helpers import CONFIGS, make_engine, make_spec

plans = [("default", "0"), ("one launch, CTA pairs", "8192")]
for name in (["cfg3", "cfg1", "tiny_gmp", "k20_ragged"] if len(sys.argv) < 2 else sys.argv[1:]):
    cfg = CONFIGS[name]
    spec = make_spec(cfg)
    x, _, eps, u = O.synthetic_batch(spec, cfg["batch"])
    for precision in ("bf16", "fp32"):
        for tag, flags in (plans if precision == "bf16" else plans[:1]):
            os.environ["GMVAE_DEBUG_FLAGS"] = flags
            eng = make_engine(cfg, precision)
            eng.initialize(3)
            for _ in range(2):
                loss = eng.train_step(x, eps=eps, gumbel_u=u)
            torch.cuda.synchronize()
            print(name, precision, tag, [round(v, 4) for v in loss.cpu().tolist()], flush=True)
            eng.close()
print("done")
