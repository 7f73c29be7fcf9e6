"""Per-kernel census of the sm_100a tensor-core / TMEM / TMA instructions in libdif_b200.so.

    python tools/sass_census.py > profiles/rNN_sass_census.md

Reads `cuobjdump -sass` of the in-tree library (no GPU needed) and counts, per kernel, the SASS mnemonics that prove
the tcgen05 path (B200_PROFILING.md): UTCHMMA / UTCQMMA (tcgen05.mma, `.2CTA` = cta_group::2), LDTM / STTM
(tcgen05.ld / st), UTMALDG (TMA tensor loads), UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus legacy HMMA (mma.sync)
which must be absent.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "deep_insight_face_b200", "libdif_b200.so")
PATTERNS = [("UTCHMMA", r"\bUTCHMMA\b"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"),
            ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("UTMAPF", r"\bUTMAPF|UTMACCTL"), ("HMMA (legacy)", r"\bHMMA\b")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    names = sorted(set(re.findall(r"Function : (\S+)", out)))
    if names:
        dm = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle.get(m.group(1), m.group(1))
            cur = re.sub(r"\(.*$", "", cur).replace("void ", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        counts[cur]["instr"] += 1
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][name] += 1
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(f"# SASS census of libdif_b200.so (cuobjdump -sass, built from {head}+, sm_100a)\n")
    print("`UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `LDTM` = tcgen05.ld, `UTMALDG` = TMA tensor load, "
          "`UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops.  Kernels without any of them (CUDA-core kernels) are listed "
          "in the second table.\n")
    cols = [n for n, _ in PATTERNS]
    print("| kernel | SASS instr | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    tot = collections.Counter()
    rest = []
    for k, c in counts.items():
        if not any(c[n] for n in cols[:5]):
            rest.append((k, c["instr"]))
            continue
        print(f"| `{k}` | {c['instr']} | " + " | ".join(str(c[n]) for n in cols) + " |")
        tot.update(c)
    print(f"| **total ({sum(1 for c in counts.values() if any(c[n] for n in cols[:5]))} tensor-core kernels)** | {tot['instr']} | "
          + " | ".join(str(tot[n]) for n in cols) + " |")
    print("\n## CUDA-core kernels (no tcgen05 / TMA instructions)\n")
    print("| kernel | SASS instr |\n|---|---|")
    for k, n in rest:
        print(f"| `{k}` | {n} |")


if __name__ == "__main__":
    sys.exit(main())
