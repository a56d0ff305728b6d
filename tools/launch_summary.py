"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (development aid)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; data = rows[hi + 1:]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); mi = hdr.index('Metric Name'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    agg.setdefault(r[ki], []).append(v)
nfw = max(1, min(len(v) for k, v in agg.items() if 'attend' in k))
tot = 0.0
for k, v in agg.items():
    if 'at::' in k:
        continue
    per = sum(v) / nfw
    tot += per
    print(f"{k.split('(')[0][:58]:58s} launches/forward {len(v)/nfw:4.1f}  {per:8.1f} us")
print(f"{'TOTAL':58s} {'':21s} {tot:8.1f} us")
