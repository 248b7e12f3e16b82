"""Runs one forward-layer-shaped tcgen05 GEMM a few times (target for `ncu -k regex:gemm_tc`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.helpers import CONFIGS, make_engine
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 2
M = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
eng = make_engine(CONFIGS["tiny_vae"], "bf16")
A = torch.randn(M, 512); B = torch.randn(512, 512)
for _ in range(3):
    eng.debug_gemm(impl, A, B, False, True)
torch.cuda.synchronize()
eng.close()
