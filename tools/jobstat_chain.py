"""Per-job timeline of the step's chained launch summed over all CTAs (test hook gmvae_debug_chain_jobstat): when every
job's first tile started and its last tile ended (globaltimer), and where the CTAs' time went.
Usage: python tools/jobstat_chain.py [batch] [workload: cfg4|cfg5] [out.json]"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gmvae_b200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
wl = sys.argv[2] if len(sys.argv) > 2 else "cfg4"
shape = dict(cfg4=dict(latent_size=64, hidden_sizes=[512, 512], mixture_components=10),
             cfg5=dict(latent_size=128, hidden_sizes=[1024, 1024], mixture_components=50))[wl]
eng = gmvae_b200.Engine("gmvae", max_batch=B, precision="bf16", seed=1, **shape)
x = (torch.rand(B, 784, device="cuda") < 0.3).to(torch.uint8)
for _ in range(5):
    eng.train_step(x)
NJ = 40
stat = torch.zeros(NJ, 8, dtype=torch.int64, device="cuda")
stat[:, 0] = -1            # ~0 as uint64
eng.lib.gmvae_debug_chain_jobstat(eng._h, stat.data_ptr())
eng.train_step(x)
torch.cuda.synchronize()
eng.lib.gmvae_debug_chain_jobstat(eng._h, None)
desc = (C.c_int * (8 * NJ))()
n = eng.lib.gmvae_debug_chain_jobs(eng._h, desc, NJ)
st = stat.cpu().tolist()
KIND = {0: "store_bf16", 1: "store_f32", 2: "bce", 3: "relumask", 4: "wgrad", 16: "rows_y_fwd", 17: "rows_z_fwd", 18: "rows_z_bwd", 19: "rows_y_bwd"}
t0 = min(s[0] for s in st[:n] if s[0] > 0)
rows = []
print(f"{'job':>3} {'kind':>11} {'M':>6} {'N':>4} {'kb':>4} {'tiles':>5} | {'start us':>8} {'end us':>8} | per tile (us): {'depwait':>7} {'mma':>6} {'accfree':>7} {'epi':>6} {'accwait':>7}")
for j in range(n):
    d = desc[8 * j:8 * j + 8]
    s = st[j]
    tiles = max(s[6], 1)
    row = dict(job=j, kind=KIND.get(d[0], str(d[0])), M=d[1], N=d[2], kblocks=d[3], tiles=d[4], splits=d[5], block_n=d[6], ndeps=d[7],
               start_us=(s[0] - t0) / 1e3, end_us=(s[1] - t0) / 1e3, depwait_us=s[2] / 1e3 / max(d[4], 1), mma_us=s[3] / 1e3 / max(d[4], 1),
               accfree_us=s[7] / 1e3 / max(d[4], 1), epi_us=s[4] / 1e3 / tiles, accwait_us=s[5] / 1e3 / tiles)
    rows.append(row)
    print(f"{j:3d} {row['kind']:>11} {d[1]:6d} {d[2]:4d} {d[3]:4d} {d[4]:5d} | {row['start_us']:8.1f} {row['end_us']:8.1f} | "
          f"{row['depwait_us']:7.2f} {row['mma_us']:6.2f} {row['accfree_us']:7.2f} {row['epi_us']:6.2f} {row['accwait_us']:7.2f}")
if len(sys.argv) > 3:
    json.dump(rows, open(sys.argv[3], "w"), indent=1)
eng.close()
