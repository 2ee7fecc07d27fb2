"""Per-kernel DRAM traffic and duration of ONE training step from an
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` capture of `bench.py --no-graph`.
Prints the per-kernel table and the totals for (a) the whole step and (b) the HSTU block stack (everything between the
embedding kernels and the head: what `roofline.traffic` in bench.py refers to)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, mi, vi, ui, idi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit"), H.index("ID")
launches = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if "byte" in u.lower():
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    elif u in ("ns", "us", "ms", "s"):
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[u]
    d = launches.setdefault(r[idi], {"name": r[ki].split("(")[0].replace("void ", "").replace("grb::", "")})
    d[r[mi]] = v
seq = list(launches.values())
starts = [i for i, d in enumerate(seq) if d["name"].startswith("embed_fwd")]
seq = seq[starts[-2]:starts[-1]] if len(starts) >= 2 else seq[starts[-1]:]
HEAD = ("tc_ce", "ce_count", "ce_finish", "embed_", "adam", "ln_fwd", "ln_bwd", "at::", "hstu_bias_index", "hstu_seq_prep", "assert_unit", "cast_flat")
agg = collections.OrderedDict()
tot = [0.0, 0.0, 0.0]
blk = [0.0, 0.0, 0.0]
n_tn = 0
for d in seq:
    rd, wr, t = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0), d.get("gpu__time_duration.sum", 0.0)
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += rd; a[2] += wr; a[3] += t
    tot[0] += rd; tot[1] += wr; tot[2] += t
    in_block = not d["name"].startswith(HEAD)
    if d["name"].startswith("tc_tn_group"):
        n_tn += 1
        in_block = n_tn > 1          # the first grouped launch of the backward pass is dE = dlogits^T x (head)
    if in_block:
        blk[0] += rd; blk[1] += wr; blk[2] += t
print(f"{'kernel':70s} {'n':>3s} {'read MB':>9s} {'write MB':>9s} {'us':>8s} {'GB/s':>7s}")
for k, (n, rd, wr, t) in sorted(agg.items(), key=lambda kv: -kv[1][3]):
    print(f"{k[:70]:70s} {n:3d} {rd/1e6:9.1f} {wr/1e6:9.1f} {t:8.1f} {(rd+wr)/t/1e3 if t else 0:7.0f}")
print(f"whole step : read {tot[0]/1e6:.0f} MB  write {tot[1]/1e6:.0f} MB  total {(tot[0]+tot[1])/1e6:.0f} MB  in {tot[2]:.0f} us (serialised)")
print(f"block stack: read {blk[0]/1e6:.0f} MB  write {blk[1]/1e6:.0f} MB  total {(blk[0]+blk[1])/1e6:.0f} MB  in {blk[2]:.0f} us (serialised)")
