"""Per-kernel SASS mnemonic counts of genrec_b200/libgenrec_b200.so (cuobjdump -sass): which kernels carry tcgen05 (UTCHMMA / LDTM),
TMA (UTMALDG / UTMASTG), mbarrier (SYNCS), legacy mma.sync (HMMA) / cp.async (LDGSTS), and generic vs shared loads/stores."""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "genrec_b200", "libgenrec_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, c, n = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        n[fn] += 1
        c[fn][m.group(1).split(".")[0]] += 1
names = subprocess.run(["c++filt"], input="\n".join(n), capture_output=True, text=True).stdout.splitlines()
dem = dict(zip(n, names))
keys = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "LDGSTS", "LDSM", "MUFU", "LDS", "STS", "LD", "ST", "LDL", "STL"]
print("SASS mnemonics per kernel (cuobjdump -sass, sm_100a).  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA tensor load/store,")
print("SYNCS = mbarrier, HMMA = mma.sync, LDGSTS = cp.async, LD/ST = generic address space, LDL/STL = local memory (spills).")
print(f"{'kernel':86s} {'instrs':>6s} " + " ".join(f"{k:>7s}" for k in keys))
for f in sorted(n, key=lambda f: dem[f]):
    d = re.sub(r"\(.*$", "", dem[f].replace("grb::", "").replace("void ", ""))
    print(f"{d[:86]:86s} {n[f]:6d} " + " ".join(f"{c[f][k]:7d}" for k in keys))
