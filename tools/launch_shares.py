"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean time, share."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(r[ui], 1.0)
    name = r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':44s} {'n':>4s} {'mean ms':>10s} {'max ms':>10s} {'share':>7s}   (cold-cache, serialised: compare shares, not absolutes;")
print(f"{'':44s} {'':4s} {'':10s} {'':10s} {'':7s}    max = a launch on the whole resident batch, the e2e pipeline launches 1/8 chunks)")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:44]:44s} {len(v):4d} {sum(v)/len(v):10.4f} {max(v):10.4f} {100*sum(v)/tot:6.1f}%")
