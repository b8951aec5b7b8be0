"""Top SASS instructions of an .ncu-rep by warp-stall samples, with instruction-class histogram:
python tools/ncu_hot.py file.ncu-rep [top_n]"""
import csv
import subprocess
import sys
from collections import Counter

f = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", f, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) <= iE:
        continue
    try:
        data.append((int(r[iN] or 0), int(r[iE] or 0), r[iS].strip(), r))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
print(f"total samples {tot}, instructions executed {sum(d[1] for d in data)}")
ops = Counter()
samp = Counter()
for n, e, s, _ in data:
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    op = op.split(".")[0]
    ops[op] += e
    samp[op] += n
print("by opcode (executed warp-instr, samples):")
for op, e in ops.most_common(22):
    print(f"  {op:12s} {e:12d} {samp[op]:8d} ({100.0 * samp[op] / max(tot, 1):5.1f} %)")
print("hottest instructions:")
for i, (n, e, s, r) in enumerate(sorted(data, key=lambda d: -d[0])[:top]):
    st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:3]
    print(f"  {n:6d} {100.0 * n / max(tot, 1):5.1f}%  x{e:<9d} {s[:70]:70s} {' '.join(f'{k}:{v}' for v, k in st if v)}")
