"""SASS evidence per kernel of libmac_b200.so: counts of the mnemonics that prove Blackwell-native code
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP).
    python profiles/sass_grep.py > profiles/r2/sass_grep.txt          (needs cuobjdump; reads the built library)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "mac_network_b200", "csrc", "libmac_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMALDG.2CTA", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "MUFU.EX2", "HMMA"]
counts = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for p in pats:
        if op == p or op.startswith(p + "."):
            counts[cur][p] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass mac_network_b200/csrc/libmac_b200.so -- per-kernel counts (kernels with none of the mnemonics omitted)")
print("# " + "  ".join(pats))
for name, pretty in zip(counts, demangle):
    c = counts[name]
    if not any(c[p] for p in pats if p not in ("SYNCS", "MUFU.EX2")):
        continue
    short = re.sub(r"\(.*", "", pretty)[:90]
    print("%-92s %s" % (short, "  ".join("%s=%d" % (p, c[p]) for p in pats if c[p])))
