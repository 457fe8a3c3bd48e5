"""Per-kernel counts of the SASS mnemonics that prove which hardware path a kernel uses, from the shipped library:
UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA tensor loads / stores), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), plus the plain tensor-core fallbacks (HMMA / IMMA) that must NOT appear.
usage: sass_summary.py [library.so] > profiles/rNN_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
    ROOT, "super-net-bayesian-image-segmentation-with-uncertainty-propagation_b200", "libsupernet_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "HMMA", "IMMA", "STG.E.ENL2.256", "LDG.E.ENL2.256"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in KEYS:
        if re.search(r"\b" + re.escape(k) + r"\b", line) or (("." in k) and k in line):
            per[cur][k] += 1
demangled = subprocess.run(["cu++filt"] + list(per), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(per, demangled)) if len(demangled) == len(per) else {k: k for k in per}


def short(n):
    n = n.replace("void sn::", "").replace("sn::", "")
    n = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", n)
    return n.replace("(int)", "").replace("(bool)", "")


tot = collections.Counter()
rows = []
for k, c in per.items():
    if not any(c[x] for x in KEYS):
        continue
    rows.append((short(names[k]), c))
    tot.update(c)
print(f"# SASS summary of `{os.path.basename(lib)}` (`cuobjdump -sass`, sm_100a): {len(per)} kernels, "
      f"{len(rows)} of them use tcgen05 / TMA / mbarrier / 256-bit global accesses\n")
print("| kernel | " + " | ".join(KEYS) + " |\n|---|" + "---|" * len(KEYS))
for n, c in sorted(rows, key=lambda r: (-r[1]["UTCHMMA"], r[0])):
    print(f"| `{n}` | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
print("| **total** | " + " | ".join(str(tot[k]) for k in KEYS) + " |")
print("\nHMMA / IMMA (mma.sync-class tensor instructions) must be 0: the tensor-core path is tcgen05 only.")
