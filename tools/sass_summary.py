#!/usr/bin/env python3
"""Per-kernel counts of the SASS opcodes that show which hardware path a kernel uses (cuobjdump -sass of the shipped library):
UTCHMMA = tcgen05.mma (".2CTA" = cta_group::2), UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, LDTM / STTM = tcgen05.ld / st,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, plus the CUDA-core floating-point and 128-bit memory instructions.

  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "laplace-dqn-snake-game_b200", "libsnake_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
OPS = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "FFMA", "HFMA2", "DFMA", "DADD", "DMUL",
       "LDG.E.128", "STG.E.128", "LDS.128", "STS.128", "SHFL", "BAR.SYNC", "BAR.ARV"]
kern, counts = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        counts[kern] = {"total": 0}
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + ".") or (o == "UTCHMMA" and op.startswith("UTCHMMA") and ".2CTA" not in op):
                if o == "UTCHMMA" and ".2CTA" in op:
                    continue
                counts[kern][o] = counts[kern].get(o, 0) + 1
                break
print("SASS opcode counts per kernel of libsnake_b200.so (sm_100a); columns with no hit are omitted")
for k in sorted(counts, key=lambda k: -counts[k]["total"]):
    c = counts[k]
    name = re.sub(r"\(.*", "", k)
    print("%-70s %6d instr  %s" % (name[:70], c["total"], "  ".join("%s=%d" % (o, c[o]) for o in OPS if o in c)))
tot = {}
for c in counts.values():
    for o, v in c.items():
        tot[o] = tot.get(o, 0) + v
print("\nwhole library: " + "  ".join("%s=%d" % (o, tot[o]) for o in ["total"] + OPS if o in tot))
