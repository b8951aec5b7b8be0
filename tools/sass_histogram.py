"""SASS opcode histogram per kernel of libdimb200.so (run where cuobjdump is installed): the Blackwell-native evidence
(UTCHMMA = tcgen05.mma, UTMALDG = TMA load, UTMAPF = TMA L2 prefetch, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, .2CTA = cta_group::2) next to the CUDA-core instruction mix.
    python tools/sass_histogram.py deep-image-matching_b200/libdimb200.so > profiles/r2_sass_histogram.txt"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict

lib = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern, cur = OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        kern[cur][m.group(1)] += 1
KEY = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "MUFU", "FFMA", "LDG", "STG", "LDS", "STS", "ATOM"]
print(f"{len(kern)} kernels in {lib}\n")
tot = Counter()
for name, c in kern.items():
    base = Counter()
    for op, n in c.items():
        root = op.split(".")[0]
        base[root] += n
        if op.startswith("UTCHMMA.2CTA") or ".2CTA" in op:
            base[root + ".2CTA"] += n
        tot[root] += n
    dm = demangle(name)
    short = re.sub(r"^void ", "", dm)
    short = re.sub(r"\(anonymous namespace\)::", "", short)
    short = (re.sub(r"\((?!anonymous).*", "", short) or name)[:150]
    print(short)
    print("    " + "  ".join(f"{k}:{base[k]}" for k in KEY if base[k]) + f"   [total {sum(c.values())}]")
print("\nlibrary totals: " + "  ".join(f"{k}:{tot[k]}" for k in KEY if tot[k]))
