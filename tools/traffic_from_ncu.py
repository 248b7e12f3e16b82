"""Sums dram__bytes_{read,write}.sum over the tcgen05 GEMM launches of one step in an `ncu --set full`
report and records it in profiles/r1_traffic.json (read by bench.py for roofline.traffic).

    python tools/traffic_from_ncu.py gpurun_out/prof_step.ncu-rep cfg4:reference:bf16
"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, key = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
def col(name): return hdr.index(name)
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
tot = 0.0; n = 0; t_us = 0.0
for d in rows[2:]:
    if len(d) != len(hdr) or not any(k in d[col("Kernel Name")] for k in ("gemm_tc_kernel", "gemm_chain_kernel")):
        continue
    tot += to_bytes(d[col("dram__bytes_read.sum")], units[col("dram__bytes_read.sum")])
    tot += to_bytes(d[col("dram__bytes_write.sum")], units[col("dram__bytes_write.sum")])
    t_us += float(d[col("gpu__time_duration.sum")].replace(",", "")) * (1e-3 if units[col("gpu__time_duration.sum")] == "ns" else 1.0)
    n += 1
path = os.path.join(ROOT, "profiles", "r1_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[key] = {"dram_bytes_tc_gemm_per_step": tot, "tc_launches_per_step": n, "ncu_us_tc_gemm_per_step": t_us, "source": os.path.basename(rep)}
json.dump(data, open(path, "w"), indent=1)
print(key, data[key])
