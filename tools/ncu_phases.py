#!/usr/bin/env python
"""Per-phase (between barriers) instruction / stall-sample shares of a kernel
from the source page of an .ncu-rep:

    ncu -i rep.ncu-rep --page source --csv > src.csv
    python tools/ncu_phases.py src.csv <pairs per launch>
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
pairs = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
ie, smp = ix['Instructions Executed'], ix['# Samples']
wf = ix['L1 Wavefronts Shared']
base = int(data[0][0], 16)
tot = sum(int(r[ie]) for r in data)
ts = sum(int(r[smp]) for r in data)
tw = sum(int(r[wf]) for r in data)
cur = None
print('phase (SASS offsets)      inst%  samples%  smem-wavefront%  static  warp-inst/pair')
for r in data:
    a = int(r[0], 16) - base
    if cur is None:
        cur = [a, 0, 0, 0, 0]
    cur[1] += int(r[ie]); cur[2] += int(r[smp]); cur[3] += 1; cur[4] += int(r[wf])
    if 'BAR' in r[1] or r is data[-1]:
        print(f'{cur[0]:#7x}..{a:#7x}  {100 * cur[1] / tot:10.1f} {100 * cur[2] / ts:9.1f} '
              f'{100 * cur[4] / max(tw, 1):12.1f} {cur[3]:10d} {cur[1] / pairs:12.0f}')
        cur = None
print('total warp-inst/pair', tot / pairs, ' smem wavefronts/pair', tw / pairs)
