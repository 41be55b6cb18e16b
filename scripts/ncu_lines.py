#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass | python scripts/ncu_lines.py [top]"""
import csv, sys, collections
rows = list(csv.reader(sys.stdin))
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cur_file = None
hdr = None
out = []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        hdr = r; continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            out.append((int(d['Instructions Executed']), int(d.get('# Samples', '0') or 0), cur_file, int(r[0]), r[1].strip()[:90]))
        except ValueError:
            pass
tot = sum(o[0] for o in out); tots = sum(o[1] for o in out)
print(f"total warp instructions {tot:.4g}, stall samples {tots}")
byfile = collections.Counter()
for o in out: byfile[o[2]] += o[0]
for f, c in byfile.most_common(): print(f"  {f:24s} {100*c/tot:5.1f}%")
for o in sorted(out, reverse=True)[:top]:
    print(f"{100*o[0]/tot:5.1f}% inst {100*o[1]/max(1,tots):5.1f}% stall  {o[2]}:{o[3]:<4d} {o[4]}")
