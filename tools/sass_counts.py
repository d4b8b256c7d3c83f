"""SASS instruction counts per kernel of the shipped library (dev tool).
    python tools/sass_counts.py > profiles/r2_sass_counts.md"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "unimp_b200", "libunimp_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKRED", "UTCBAR", "UTCBAR.MULTICAST", "UCGABAR_ARV",
        "MUFU.EX2", "BRA.U.ANY", "HMMA"]
counts = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for p in pats:
        if p == "UTCBAR":
            if op.startswith("UTCBAR") and "MULTICAST" not in op:
                counts[cur][p] += 1
        elif p == "HMMA":
            if op.startswith("HMMA"):
                counts[cur][p] += 1
        elif op.startswith(p) or (p in ("UTCBAR.MULTICAST",) and op.startswith("UTCBAR") and "MULTICAST" in op):
            counts[cur][p] += 1
print("# SASS instruction counts of the shipped library (round 2)\n")
print("`cuobjdump -sass unimp_b200/libunimp_b200.so`, per kernel (`tools/sass_counts.py`); the mnemonics of "
      "B200_PROFILING.md: `UTCHMMA` = tcgen05.mma, `LDTM` / `STTM` = tcgen05.ld / st, `UTMALDG` / `UTMASTG` / `UBLKRED` = "
      "TMA load / store / bulk reduce, `UTCBAR` = tcgen05.commit (`.MULTICAST`: to all CTAs of the cluster), `UCGABAR_ARV` "
      "= cluster barrier.  `BRA.U.ANY` (the waterfall loops of round 1, DESIGN.md §4.2) must be 0 everywhere.  Warp-level "
      "`HMMA` (`mma.sync`) appears in exactly one place, `linear_small_m_kernel` — the decode step's weight-streaming "
      "projection for <= 8 rows, an HBM-bound kernel whose fragments are filled straight from global memory "
      "(DESIGN.md §8); every contraction of the training path is `UTCHMMA`.\n")
print("| kernel | " + " | ".join(pats) + " |")
print("|---|" + "---:|" * len(pats))
for k, c in counts.items():
    if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"] or c["HMMA"]:
        print(f"| `{k}` | " + " | ".join(str(c[p]) for p in pats) + " |")
