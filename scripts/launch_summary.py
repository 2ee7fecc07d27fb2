"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share of one step."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value'); ui = H.index('Metric Unit'); mi = H.index('Metric Name')
seq = []
for r in rows[hdr + 1:]:
    if len(r) > vi and r[mi] == 'gpu__time_duration.sum':
        v = float(r[vi].replace(',', '')); u = r[ui]
        v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
        seq.append((r[ki].split('(')[0].replace('void ', '').replace('grb::', ''), v))
# one training step = from one embed_fwd_kernel to the next
starts = [i for i, (k, _) in enumerate(seq) if k.startswith('embed_fwd')]
if len(starts) >= 2: seq = seq[starts[-2]:starts[-1]]
agg = collections.OrderedDict()
for k, v in seq:
    c, t = agg.get(k, (0, 0.0)); agg[k] = (c + 1, t + v)
tot = sum(t for _, t in agg.values())
print(f"one step: {len(seq)} launches, {tot:.0f} us (serialised, cold cache)")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us {100*t/tot:5.1f}%  x{c:<3d} {k[:100]}")
