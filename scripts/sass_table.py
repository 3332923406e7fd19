#!/usr/bin/env python
"""SASS mnemonic counts per kernel of libkvquant_b200.so (cuobjdump -sass): the evidence that the hot kernels use
TMA (UTMALDG / UBLKCP), mbarriers (SYNCS), packed fp32 FMA (FFMA2), the mixed-precision FMA (FHFMA) and REDUX."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "kvquant_b200/libkvquant_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = ["UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FHFMA", "FFMA", "LDS", "LDG", "PRMT", "REDUX", "RED", "ATOMS", "ATOMG", "SHFL", "BAR"]
cur = None
cnt = collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("kvq::", "")
        cnt[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(2)
        cnt[cur][op] += 1
        cnt[cur]["_total"] += 1
print("%-46s %6s " % ("kernel", "instr") + " ".join("%7s" % w for w in want))
for k, c in cnt.items():
    if c["_total"] < 60:
        continue
    print("%-46s %6d " % (k[:46], c["_total"]) + " ".join("%7d" % c[w] for w in want))
