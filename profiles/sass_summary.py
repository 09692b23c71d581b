"""Per-kernel SASS mnemonic counts of libipsr_sm100.so (cuobjdump -sass): evidence that the shipped binary holds tcgen05
(UTCHMMA / UTCBAR / LDTM = tcgen05.mma / commit / ld), TMA bulk copies (UBLKCP) and mbarriers (SYNCS), kernel by kernel.
usage: python profiles/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "deepinpainting_b200", "lib", "libipsr_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UBLKCP", "SYNCS", "UCGABAR", "SHFL", "ATOMS", "ATOMG", "RED", "LDS", "STS", "LDG", "STG", "FFMA", "HFMA2"]
cur, counts, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if cur:
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["total"] += 1
            for w in WATCH:
                if op.startswith(w):
                    counts[cur][w] += 1
print("library:", os.path.relpath(so, ROOT), " arch:", ", ".join(sorted(arch)))
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
print("whole library:", ", ".join("%s %d" % (w, tot[w]) for w in WATCH if tot[w]))
print()
print("%-78s %7s  %s" % ("kernel", "instrs", "mnemonics of interest"))
for k, c in counts.items():
    print("%-78s %7d  %s" % (k[:78], c["total"], ", ".join("%s %d" % (w, c[w]) for w in WATCH if c[w])))
