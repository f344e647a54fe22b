"""Per-kernel census of the Blackwell-specific SASS in the built library (tcgen05 MMA, TMA, TMEM loads, tcgen05 commits, mbarrier ops):
    python tools/sass_census.py [path/to/libskit_b200.so] > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "visual-tactile-synthesis_b200", "csrc", "libskit_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ops = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "UTCATOMSWS", "REDG", "ATOMG")
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    if cur:
        for op in ops:
            if re.search(r"\b%s\b" % op, line):
                counts[cur][op] += 1
print("%-64s " % "kernel" + " ".join("%8s" % o for o in ops))
tot = collections.Counter()
for k, c in counts.items():
    if any(c[o] for o in ops[:6]):
        print("%-64s " % k[:64] + " ".join("%8d" % c[o] for o in ops))
    tot.update(c)
print("%-64s " % ("TOTAL over %d kernels" % len(counts)) + " ".join("%8d" % tot[o] for o in ops))
