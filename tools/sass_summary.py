"""Per-kernel static counts of the tcgen05 / TMA / mbarrier instructions in the built library (cuobjdump -sass).
Usage: python tools/sass_summary.py [out.md]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gmvae_b200", "libgmvae_b200.so")
COLS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "UTCBAR.2CTA.MULTICAST", "LDTM", "UTMALDG", "UTMALDG.2CTA", "UTMALDG.MULTICAST.2CTA", "UTMASTG", "UTMAREDG",
        "UTCATOMSWS", "UCGABAR", "ELECT", "MUFU"]


def classify(op):
    if op.startswith("UTCHMMA"): return "UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"
    if op.startswith("UTCBAR"): return "UTCBAR.2CTA.MULTICAST" if ".2CTA" in op else "UTCBAR"
    if op.startswith("LDTM"): return "LDTM"
    if op.startswith("UTMALDG"):
        return "UTMALDG.MULTICAST.2CTA" if ".MULTICAST" in op else "UTMALDG.2CTA" if ".2CTA" in op else "UTMALDG"
    for k in ("UTMASTG", "UTMAREDG", "UTCATOMSWS", "UCGABAR", "ELECT", "MUFU"):
        if op.startswith(k): return k
    return None


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    dem = {}
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); counts[cur] = collections.Counter(); continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            counts[cur]["_n"] += 1
            c = classify(m.group(1))
            if c: counts[cur][c] += 1
    names = list(counts)
    d = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    rows = []
    for n, dn in zip(names, d):
        c = counts[n]
        if not any(c[k] for k in COLS[:11]): continue
        dn = re.sub(r"^void ", "", dn); dn = re.sub(r"\(.*$", "", dn).replace("gmvae::", "").replace("__nv_bfloat16", "bf16")
        rows.append((dn, c))
    rows.sort(key=lambda r: (-("gemm_chain" in r[0]), r[0]))
    lines = ["# SASS summary of `gmvae_b200/libgmvae_b200.so`", "",
             "`python tools/sass_summary.py` (`cuobjdump -sass`, static instruction counts per kernel, sm_100a). `UTCHMMA` = tcgen05.mma (`.2CTA` = "
             "cta_group::2), `UTCBAR` = tcgen05.commit (`.2CTA.MULTICAST` = multicast to the CTAs of a pair / cluster), `LDTM` = tcgen05.ld, `UTMALDG` / "
             "`UTMASTG` / `UTMAREDG` = TMA load / store / reduce-add (`.2CTA` = bytes counted on the pair leader's barrier, `.MULTICAST.2CTA` = the "
             "box lands in several CTAs of the cluster: the opt-in quad mode), `UTCATOMSWS` = tcgen05.alloc, `UCGABAR` = cluster barrier, `ELECT` = "
             "elect.sync (the warp-uniform producer / MMA roles issue through one elected lane).", "",
             "| kernel | SASS instr | " + " | ".join(COLS) + " |", "|---|---:|" + "---:|" * len(COLS)]
    for dn, c in rows:
        lines.append(f"| `{dn}` | {c['_n']} | " + " | ".join(str(c[k]) for k in COLS) + " |")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main()
