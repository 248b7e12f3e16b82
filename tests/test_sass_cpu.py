"""Static checks on the SASS of the built library (no GPU needed: cuobjdump reads the cubin).  They pin what the round's profiling
found: the chained kernel's producer and MMA roles must issue their TMA / tcgen05 instructions through ONE elected lane of a warp that
runs with uniform control flow.  Run by a single thread inside `if (lane == 0)` the compiler wrapped every UTMALDG / UTCHMMA in a vote
loop (ELECT ... BRA.U.ANY) behind R2UR chains -- ~130 dependent instructions per k-block, which paced the whole main loop
(profiles/r2_ablation_loads.md)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gmvae_b200", "libgmvae_b200.so")


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    from gmvae_b200 import build
    build.build()
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append(m.group(1).strip())
    return funcs


def chain(funcs, cl):
    names = [n for n in funcs if "gemm_chain_kernel" in n and n.endswith(f"EEELi{cl}EEEvT_") and "Li40ELi112" in n]
    assert len(names) == 1, names
    return funcs[names[0]]


def opcode(ins):
    return re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0]


@pytest.mark.parametrize("cl", [1, 2, 4])
def test_chain_kernel_issues_through_an_elected_lane(sass, cl):
    ins = chain(sass, cl)
    ops = [opcode(i) for i in ins]
    assert any(o.startswith("ELECT") for o in ops)
    # no vote loop around a TMA load or an MMA: the instruction right after the issue is never the loop's back edge test
    # (one exception: the epilogue's legacy TMA load of a ReLU-mask source box, issued by lane 0 of an epilogue warp)
    looped = [k for k, o in enumerate(ops) if (o.startswith("UTMALDG") or o.startswith("UTCHMMA"))
              and any(w.startswith("BRA.U.ANY") for w in ops[k + 1:k + 4])]
    assert len(looped) <= 1 and all(ops[k].startswith("UTMALDG.2D ") or ops[k] == "UTMALDG.2D" for k in looped), (cl, [ins[k] for k in looped])
    assert sum(o.startswith("UTMALDG") for o in ops) >= 8
    # the four tcgen05.mma of a k-block are issued back to back
    idx = [k for k, o in enumerate(ops) if o.startswith("UTCHMMA")]
    assert len(idx) >= 4 and idx[3] - idx[0] == 3, idx


def test_pair_and_quad_forms(sass):
    pair = [opcode(i) for i in chain(sass, 2)]
    assert sum(o.startswith("UTCHMMA.2CTA") for o in pair) >= 4
    assert any(o.startswith("UTCBAR.2CTA.MULTICAST") for o in pair)
    assert any(o.startswith("UTMALDG.2D.2CTA") for o in pair)
    assert any(o.startswith("UTMASTG") for o in pair) and any(o.startswith("UTMAREDG") for o in pair)
    quad = [opcode(i) for i in chain(sass, 4)]
    assert any(o.startswith("UTMALDG.2D.MULTICAST.2CTA") for o in quad)
    single = [opcode(i) for i in chain(sass, 1)]
    assert not any(".2CTA" in o for o in single)


def test_one_gemm_kernel_issues_through_an_elected_lane(sass):
    names = [n for n in sass if "gemm_tc_kernel" in n]
    assert names
    for n in names:
        ops = [opcode(i) for i in sass[n]]
        assert any(o.startswith("ELECT") for o in ops), n
        for k, o in enumerate(ops):
            if o.startswith("UTMALDG") or o.startswith("UTCHMMA"):
                assert not any(w.startswith("BRA.U.ANY") for w in ops[k + 1:k + 4]), (n, k)
