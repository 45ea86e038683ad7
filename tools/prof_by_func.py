#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump per device function (symbol ranges from cuobjdump -elf)."""
import csv, re, sys
symfile, srcfile = sys.argv[1], sys.argv[2]
syms = []
for l in open(symfile):
    a, sz, name = l.split()
    a, sz = int(a, 16), int(sz, 16)
    if sz and "$" in name:
        syms.append((a, a + sz, re.sub(r"^_ZN3grs\d+", "", name.split("$")[-1])[:18]))
rows = list(csv.reader(open(srcfile)))
h = rows[1]
cols = ["Instructions Executed", "# Samples", "stall_no_inst", "stall_barrier", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_branch_resolving", "stall_selected", "stall_not_selected"]
idx = [h.index(c) for c in cols]
ai = h.index("Address")
base, agg = None, {}
for r in rows[2:]:
    try:
        a = int(r[ai], 16)
    except ValueError:
        continue
    if base is None:
        base = a
    off, nm = a - base, "main"
    for lo, hi, n in syms:
        if lo <= off < hi:
            nm = n
            break
    v = agg.setdefault(nm, [0] * len(cols))
    for k, i in enumerate(idx):
        v[k] += int(r[i])
tot, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print("%-20s %6s %6s | %% of its samples: %5s %5s %5s %5s %5s %5s %5s %5s" % ("function", "inst%", "samp%", "noins", "barr", "wait", "shsb", "lgsb", "brres", "sel", "notsel"))
for n, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    s = max(v[1], 1)
    print("%-20s %6.1f %6.1f | %23.0f %5.0f %5.0f %5.0f %5.0f %5.0f %5.0f %5.0f" % ((n, 100 * v[0] / tot, 100 * v[1] / ts) + tuple(100 * x / s for x in v[2:])))
print("total warp instructions %.3e, samples %d" % (tot, ts))
