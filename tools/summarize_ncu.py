"""Turns ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1_ncu_launches.md "command line"
    python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r1_ncu_gemm_full.md
"""
import collections
import csv
import re
import subprocess
import sys


def short_name(name):
    m = re.search(r"gemm_tc_kernel<\(int\)(\d+), \(bool\)(\d), \(bool\)(\d), gmvae::(\w+)(<[^>]*>)?", name)
    if m:
        major = "MN/MN (wgrad)" if m.group(2) == "1" else "K/K"
        return f"gemm_tc_kernel<BN={m.group(1)}, {major}, {m.group(4)}{(m.group(5) or '').replace('__nv_bfloat16', 'bf16')}>"
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("gmvae::", "").replace("__nv_bfloat16", "bf16")[:70]


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v


def launches(src, dst, cmd):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
        d[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
    items = []
    for d in per.values():
        t = to_us(*d["gpu__time_duration.sum"])
        tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", ("", ""))[0]
        items.append((short_name(d["name"]), d["grid"], t, tp))
    starts = [i for i, it in enumerate(items) if it[0].startswith(("convert_x", "prologue_kernel"))]
    step = items[starts[-2]:starts[-1]] if len(starts) >= 2 else items
    foreign = [i for i in step if i[0].startswith("at::")]           # bench.py's untimed L2 flush between steps (torch add_)
    step = [i for i in step if not i[0].startswith("at::")]
    tot = sum(i[2] for i in step)
    agg = collections.OrderedDict()
    for n, g, t, tp in step:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1; a[1] += t
    with open(dst, "w") as f:
        f.write(f"# ncu launch list of one training step\n\nCommand: `{cmd}`\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active... --clock-control none`; per-launch times are cold-cache and "
                "serialised, so the SHARE of the step is what is comparable with the CUDA-event numbers of bench.py, not the absolute.\n\n")
        f.write(f"One step = {len(step)} launches, {tot:.1f} us summed under ncu"
                + (f" ({len(foreign)} torch launch(es) between the steps -- bench.py's untimed L2 flush -- left out)" if foreign else "") + ".\n\n## By kernel\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{n}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |\n")
        f.write("\n## In launch order\n\n| # | kernel | grid | us | tensor pipe active % |\n|---:|---|---|---:|---:|\n")
        for i, (n, g, t, tp) in enumerate(step):
            f.write(f"| {i} | `{n}` | {g} | {t:.1f} | {tp[:5]} |\n")
    print("wrote", dst, len(step), "launches", round(tot, 1), "us")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum"]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full capture: `{src}`\n\n| launch | " + " | ".join(k.split(".")[0] for k in keys) + " |\n|---|" + "---:|" * len(keys) + "\n")
        for d in rows[2:]:
            if len(d) != len(hdr):
                continue
            name = short_name(d[hdr.index("Kernel Name")])
            f.write(f"| `{name}` | " + " | ".join((d[hdr.index(k)] + " " + units[hdr.index(k)]) if k in hdr else "-" for k in keys) + " |\n")
        src_csv = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--launch-count", "1"], capture_output=True, text=True).stdout
        srows = list(csv.reader(src_csv.splitlines()))
        if len(srows) > 2:
            h = srows[1]
            ix = {k: i for i, k in enumerate(h)}
            data, seen = [], set()
            for r in srows[2:]:
                if len(r) == len(h) and r[0] != "Address" and r[0] not in seen:
                    seen.add(r[0]); data.append(r)
            def I(v):
                try: return int(v)
                except ValueError: return 0
            tot = sum(I(r[ix["# Samples"]]) for r in data) or 1
            stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
            agg = sorted(((sum(I(r[ix[k]]) for r in data), k) for k in stalls), reverse=True)[:6]
            f.write(f"\n## Warp-stall samples of the first captured launch ({tot} samples)\n\n")
            f.write(", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for v, k in agg) + "\n\n| samples | share | SASS | top stall |\n|---:|---:|---|---|\n")
            for r in sorted(data, key=lambda r: -I(r[ix["# Samples"]]))[:14]:
                s = I(r[ix["# Samples"]])
                top = max(stalls, key=lambda k: I(r[ix[k]]))
                f.write(f"| {s} | {100 * s / tot:.1f}% | `{r[ix['Source']][:70]}` | {top[6:]} |\n")
    print("wrote", dst)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])
