"""Per-tile clock64 timeline of one CTA through the chained-GEMM launches of a cfg4 training step
(test hook gmvae_debug_chain_trace).  Usage: python tools/trace_chain.py [cta] [config]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.helpers import CONFIGS, make_engine

cta = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cfg = dict(CONFIGS["cfg3"], batch=int(sys.argv[2]) if len(sys.argv) > 2 else 16384)   # cfg4 = cfg3 shapes at batch 16384
eng = make_engine(cfg, "bf16")
eng.initialize(2024)
B = cfg["batch"]
x = (torch.rand(B, 784, device="cuda") < 0.3).to(torch.uint8)
for _ in range(3):
    eng.train_step(x)
trace = torch.zeros(8 * 64 * 16, dtype=torch.int64, device="cuda")
eng.lib.gmvae_debug_chain_trace(eng._h, trace.data_ptr(), cta)
eng.train_step(x)
torch.cuda.synchronize()
eng.lib.gmvae_debug_chain_trace(eng._h, None, 0)
tr = trace.cpu().view(8, 64, 16)
names = {0: "p_dep0", 1: "p_dep1", 2: "p_done", 3: "m_start", 4: "m_full0", 5: "m_done", 6: "e_start", 7: "e_pre", 8: "e_acc", 9: "e_rows", 10: "e_sig"}
for L in range(8):
    if int(tr[L].abs().sum()) == 0:
        continue
    t0 = min(int(v) for v in tr[L, :, [0, 3, 6]].flatten() if int(v) > 0)
    print(f"== chain launch {L} (CTA {cta}); cycles since the CTA's first stamp")
    last_end = 0
    for it in range(64):
        row = tr[L, it]
        if int(row[6]) == 0 and int(row[3]) == 0:
            break
        g = lambda i: int(row[i]) - t0 if int(row[i]) else -1
        print(f"  tile {it:2d} job {int(row[15])} local {int(row[14]):5d} | prod dep {g(0):7d}->{g(1):7d} (wait {g(1)-g(0):6d}) issued {g(2):7d} | "
              f"mma start {g(3):7d} first {g(4):7d} done-issue {g(5):7d} | epi start {g(6):7d} pre {g(7):7d} acc {g(8):7d} rows {g(9):7d} sig {g(10):7d} "
              f"(epi {g(10)-g(8):6d})")
    print(f"   total {max(int(v) for v in tr[L].flatten() if int(v) > 1000) - t0} cycles")
eng.close()
