#!/usr/bin/env python
"""Census of the Blackwell-specific SASS mnemonics per kernel of the built library (cuobjdump -sass; no GPU needed):
   python tools/sass_census.py > profiles/r01_sass_census.md"""
import collections
import re
import subprocess

KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.MULTICAST", "UTCBAR", "LDTM", "UTCATOMSWS", "SYNCS", "HMMA", "LDSM", "LDGSTS",
        "ACQBULK", "PREEXIT", "REDG", "ATOMG", "MUFU", "SHFL"]
out = subprocess.run(["cuobjdump", "-sass", "paella_b200/libpaella_b200.so"], capture_output=True, text=True).stdout
fn, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = fn.replace("(anonymous namespace)::", "")
        fn = re.sub(r"\(.*", "", fn).replace("void ", "").replace("pb::", "")
        fn = fn.replace("(int)", "").replace("(bool)", "")
        counts[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        counts[fn]["_total"] += 1
        base = op.split(".")[0]
        if base in KEYS:
            counts[fn][base] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            counts[fn]["UTCHMMA.2CTA"] += 1
        if base == "UTMALDG" and "MULTICAST" in op:
            counts[fn]["UTMALDG.MULTICAST"] += 1
print("# SASS census of `paella_b200/libpaella_b200.so` (sm_100a)\n")
print("`cuobjdump -sass`, static instruction counts per kernel.  UTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2), UTMALDG = TMA")
print("tensor load, LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, ACQBULK / PREEXIT =")
print("griddepcontrol.wait / launch_dependents, HMMA/LDSM/LDGSTS = mma.sync / ldmatrix / cp.async (attention core).\n")
cols = [k for k in KEYS if any(c[k] for c in counts.values())]
print("| kernel | instr | " + " | ".join(cols) + " |")
print("|---|---:|" + "---:|" * len(cols))
groups = collections.OrderedDict()
for fn, c in counts.items():
    groups.setdefault(fn, c)
for fn, c in groups.items():
    if c["_total"] < 40:
        continue
    print(f"| `{fn[:90]}` | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")
