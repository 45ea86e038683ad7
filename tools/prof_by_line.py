#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS dump with `nvdisasm -g -c` line info: warp-stall samples and executed instructions
per CUDA source line.  Usage: prof_by_line.py dis.txt kernel_section_substring src.csv [top]"""
import csv, re, sys, collections
dis, key, srcfile = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
line_of, cur, on = {}, None, False
for l in open(dis):
    if l.startswith("//---") and ".text." in l:
        on = key in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(srcfile)))
h = rows[1]
ai, si, ii = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
base, agg = None, collections.defaultdict(lambda: [0, 0])
for r in rows[2:]:
    try:
        a = int(r[ai], 16)
    except ValueError:
        continue
    if base is None:
        base = a
    k = line_of.get(a - base, ("?", 0))
    agg[k][0] += int(r[si]); agg[k][1] += int(r[ii])
ts, ti = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
srcs = {}
for (f, n), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    if f not in srcs:
        try:
            srcs[f] = open("/root/repo/mujoco_rl_manipulate_unknown_objects_b200/csrc/" + f).read().splitlines()
        except OSError:
            srcs[f] = []
    text = srcs[f][n - 1].strip()[:110] if 0 < n <= len(srcs[f]) else ""
    print("%5.2f%% samp %5.2f%% inst  %-20s:%4d  %s" % (100 * v[0] / ts, 100 * v[1] / ti, f, n, text))
