#!/usr/bin/env python3
"""Aggregates the ncu source-page export of k_rollout_ws (tools/ncu_rollout_source.sh -> gpurun_out/ro_source.csv) by code region
and by how many warps of a CTA execute an instruction per step (1 = a single-warp role, N = the expansion warps):
  python tools/ncu_rollout_stalls.py gpurun_out/ro_source.csv > profiles/r02_ncu_rollout_ws_stalls.txt"""
import collections
import csv
import sys

T, CTAS = 200, 256
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
data = [r for r in rows[2:] if len(r) == len(h)]
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
src = [r[ix["Source"]].strip() for r in data]
syncs = [i for i, x in enumerate(src) if "BAR.SYNC" in x]
arv = [i for i, x in enumerate(src) if x.startswith("BAR.ARV")]
# barrier order in the SASS: init __syncthreads, logic EMPTY, expansion FULL, XB, XB ..., mask FULL ..., final __syncthreads
lo0, lo1 = syncs[1] + 1, arv[0] + 1
b0, b1 = syncs[2] + 1, syncs[3] + 1
c0, c1 = syncs[3] + 1, syncs[4] + 1
d0, d1 = syncs[4] + 1, syncs[-1] + 1
U = T * CTAS
xw = max(round(int(r[ix["Instructions Executed"]] or 0) / U) for r in data[c0:c1])


def agg(a, b, cls):
    by, ex = collections.Counter(), 0.0
    for r in data[a:b]:
        e = int(r[ix["Instructions Executed"]] or 0) / U
        if round(e) != cls:
            continue
        ex += e
        for k in stalls:
            v = int(r[ix[k]] or 0)
            if v:
                by[k[6:]] += v
    tot = max(sum(by.values()), 1)
    return round(ex / cls), tot, ", ".join("%s %d%%" % (k, round(100 * v / tot)) for k, v in by.most_common(6))


print("ncu --set full --import-source on, %s at config 2 (4,096 envs x 200 steps): warp-stall sampling per SASS instruction" % rows[0][1])
print("(tools/ncu_rollout_source.sh), aggregated by code region and by the number of warps of a CTA that execute the instruction per step")
print("(1 = a single-warp role, %d = the expansion warps).  instr/step is per warp.\n" % xw)
for name, a, b, cls in (("logic warp: loop body after its EMPTY barrier .. FULL arrive", lo0, lo1, 1),
                        ("two role warps: board conversion (to_full / unit bytes / board_planes)", b0, b1, 1),
                        ("expansion warps: FULL wait + loop head", b0, b1, xw),
                        ("expansion warps: expand_obs + the barriers among them", c0, c1, xw),
                        ("mask warp: losing_mask3 + scalar stores + its FULL wait", d0, d1, 1),
                        ("expansion warps: tail", d0, d1, xw)):
    n, tot, txt = agg(a, b, cls)
    print("%-78s instr/step %4d  samples %5d  %s" % (name, n, tot, txt))
print("\nReading: 'barrier' on the logic warp is its wait at EMPTY (the consumers set the pace).  The clock64 counters of tools/rollout_probe.py")
print("cannot see such waits: after BAR.SYNC.DEFER_BLOCKING the clock read issues before the warp blocks, so they book the wait as work.")
