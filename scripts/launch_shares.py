#!/usr/bin/env python
"""Per-kernel share of device time from an ncu launch list (gpu__time_duration.sum CSV)."""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, mi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[mi] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '')
    agg.setdefault(name, []).append(float(r[vi].replace(',', '')))
tot = sum(sum(v) for v in agg.values())
print("kernel,launches,total_ms,avg_us,share")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%s,%d,%.3f,%.1f,%.4f" % (k, len(v), sum(v) / 1e6, sum(v) / len(v) / 1e3, sum(v) / tot))
print("TOTAL,%d,%.3f,," % (sum(len(v) for v in agg.values()), tot / 1e6))
