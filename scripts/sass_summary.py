#!/usr/bin/env python
"""profiles/sass_tcgen05.txt: per kernel of libgbdr.so, how often the SASS uses the Blackwell units the design claims
(cuobjdump -sass; B200_PROFILING.md names the mnemonics): UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st
(TMEM), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA engine, 1-D), UTMALDG = TMA tensor-map load, SYNCS = mbarrier,
ATOMS = shared-memory atomics, REDUX / MATCH / VOTE = the warp collectives of the beam kernel."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "gbnns_dim_red_b200/libgbdr.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
MN = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "ATOMS", "REDUX", "MATCH", "VOTE", "LDGSTS"]
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"gbdr::\(anonymous namespace\)::|\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(.*", "", cur)
        cur = re.sub(r"^void ", "", cur).replace("gbdr::", "")
        counts[cur] = collections.Counter()
        continue
    if cur:
        for mn in MN:
            if re.search(r"\b" + mn + r"\b|\b" + mn + r"\.", line):
                counts[cur][mn] += 1
print(f"cuobjdump -sass {lib} | per-kernel counts of {', '.join(MN)} (kernels with none of them omitted)")
agg = collections.OrderedDict()
for k, c in counts.items():
    if not c:
        continue
    base = re.sub(r"<.*", "", k)
    if base.startswith("beam_search_v2_kernel"):
        agg.setdefault(base, []).append(c)
    else:
        print(f"{k[:110]:110s} " + " ".join(f"{mn}={c[mn]}" for mn in MN if c[mn]))
for k, cs in agg.items():
    k = f"{k}<{len(cs)} instantiations: list capacity x row width x visited format; min..max per kernel>"
    print(f"{k:110s} " + " ".join(f"{mn}={min(c[mn] for c in cs)}..{max(c[mn] for c in cs)}" for mn in MN if any(c[mn] for c in cs)))
