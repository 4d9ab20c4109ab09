#!/usr/bin/env python
"""Attribute the stall samples / executed instructions of an ncu capture of k_wres_chunk to the inlined device functions of
qp_wres.cuh.  usage: ncu_regions.py src2.csv (from --page source --print-source cuda,sass --csv) src.csv (--page source --csv)"""
import csv, sys, bisect
src2, src = sys.argv[1], sys.argv[2]
bounds = [(42, 'helpers'), (73, 'gload/gapply (Q, Q^-1 products)'), (106, 'w_gv (G v)'), (129, 'w_gtu (G^T u)'), (154, 'lane'), (166, 'helpers'),
          (185, 'w_diag'), (224, 'fix_tiles'), (289, 'w_factor (TRSM + trailing DMMA)'), (334, 'w_fwd'), (359, 'w_bwd'), (393, 'w_pieces'),
          (426, 'kernel body')]
if len(sys.argv) > 3:
    bounds = eval(open(sys.argv[3]).read())
def region(line):
    r = 'pre'
    for b, name in bounds:
        if line >= b: r = name
    return r
rows = list(csv.reader(open(src2)))
addr2 = {}
cur = None; cur_file = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] in ('Line No', 'Function Name'): continue
    if len(r) > 2 and r[2] == '-':
        try: cur = (cur_file, int(r[0]))
        except ValueError: cur = None
        continue
    if len(r) > 2 and r[2].startswith('0x') and cur:
        addr2[r[2]] = cur
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, ismp, iex = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stn = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {}
last = 'pre'
tot_s = tot_e = 0
for r in rows[2:]:
    if len(r) <= iex: continue
    loc = addr2.get(r[ia])
    if loc and loc[0] == 'qp_wres.cuh' and loc[1] >= 66:
        last = region(loc[1])
    elif loc and loc[0] == 'qp_wres.cuh':
        pass  # shuffle / reduction helpers: keep the enclosing region
    a = agg.setdefault(last, dict(s=0.0, e=0.0, st={}))
    s = float(r[ismp] or 0); e = float(r[iex] or 0)
    a['s'] += s; a['e'] += e; tot_s += s; tot_e += e
    for h in stn:
        v = float(r[hdr.index(h)] or 0)
        if v: a['st'][h[6:]] = a['st'].get(h[6:], 0) + v
print(f"total samples {tot_s:.0f}, executed warp instructions {tot_e:.0f}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['s']):
    top = sorted(a['st'].items(), key=lambda kv: -kv[1])[:4]
    print(f"{100*a['s']/tot_s:5.1f}% samples {100*a['e']/tot_e:5.1f}% instr  {k:40s} " + ', '.join(f"{n} {100*v/a['s']:.0f}%" for n, v in top))
