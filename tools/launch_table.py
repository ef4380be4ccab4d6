#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log: the last N launches."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
last_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[0]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit'); ii = hdr.index('ID')
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[ii], {'k': r[ki].split('(')[0][-44:]})
    v = float(r[vi].replace(',', '')); u = r[ui]
    if r[mi].startswith('gpu__time'):
        d['us'] = v / 1e3 if u == 'ns' else (v if u == 'us' else v * 1e3)
    else:
        d['rd' if 'read' in r[mi] else 'wr'] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
last = list(per.values())[-last_n:]
tot = sum(d['us'] for d in last)
for d in last:
    rd, wr = d.get('rd', 0), d.get('wr', 0)
    print(f"{d['k']:46s} {d['us']:9.1f} us {d['us']/tot*100:5.1f}%  rd {rd/1e6:8.1f} MB wr {wr/1e6:8.1f} MB  {(rd+wr)/max(d['us'],1e-9)/1e3:7.0f} GB/s")
print('total', round(tot, 1), 'us')
