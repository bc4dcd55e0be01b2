#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): python profiles/launch_table.py file.csv [steps]"""
import collections, csv, sys
def table(path, steps=1):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr, start = r, i + 1
            break
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) <= vi: continue
        name = r[ki].split('(')[0][:44]
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, a in agg.items():
        print(f"{k:46s} n={a[0]:4d} total/step={a[1]/1e3/steps:9.3f} ms mean={a[1]/a[0]:9.1f} us share={100*a[1]/tot:5.1f}%")
    print(f"{'all':46s}        total/step={tot/1e3/steps:9.3f} ms")
if __name__ == '__main__':
    table(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
