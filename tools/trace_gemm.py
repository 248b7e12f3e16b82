"""Per-phase clock64 timeline of CTA 0 of the tcgen05 GEMM kernel (test hook gmvae_debug_gemm impl>=2:
2 = forward layer epilogue (bias+ReLU+bf16), 3 = ReLU-mask epilogue, 4 = ReLU-mask + fused bias gradient)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.helpers import CONFIGS, make_engine

eng = make_engine(CONFIGS["tiny_vae"], "bf16")
names = ["prod_start", "prod_end", "mma_start", "mma_tmem_free", "mma_first_full", "mma_done", "epi_start", "epi_tmem_full",
         "chunk0", "chunk1", "chunk2", "chunk3", "epi_end", "ld0", "ld1", "ld2"]
shapes = [(16384, 512, 512)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for impl, label in ((2, "EpiStore<bf16> relu+bias"), (3, "EpiReluMask"), (4, "EpiReluMask+colsum")):
    for (M, N, K) in shapes:
        A = torch.randn(M, K); B = torch.randn(N, K)
        for rep in range(2):
            out = eng.debug_gemm(impl, A, B, False, True)
        tr = out.view(-1)[: 4096 * 2].view(torch.int64).cpu().view(-1, 16)
        t0 = int(tr[0, 0])
        print(f"== {label} M={M} N={N} K={K}: cycles relative to the producer's first issue (CTA 0)")
        for it in range(4):
            row = tr[it]
            if int(row[0]) == 0 and it > 0:
                break
            print("  tile", it, " ".join(f"{n}={int(row[i]) - t0}" for i, n in enumerate(names) if int(row[i]) != 0))
eng.close()
