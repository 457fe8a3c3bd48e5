"""Who waits for whom: every mbarrier wait of a kernel in an ncu report, with its first-try count, retry count and
stall samples, attributed to CUDA source lines (same join as ncu_by_line.py).
usage: ncu_waits.py <report.ncu-rep> <library.so>"""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, lib = sys.argv[1:3]
# a report, or the csv of its source page (`ncu -i rep --page source --csv`, made on the GPU box)
raw = open(rep).read() if rep.endswith(".csv") else \
    subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
kernel = rows[0][1]
h = rows[1]
ia, ii, isamp, isrc = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
sass = [(int(r[ia], 16), int(r[ii] or 0), int(r[isamp] or 0), r[isrc]) for r in rows[2:] if len(r) >= len(h)]
base = sass[0][0]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
m = re.match(r"void sn::(\w+)<(.*)>\(", kernel)
fn = m.group(1)
enc = "".join(("Lb%sE" % a.split(")")[-1]) if "bool" in a else ("Li%sE" % a.split(")")[-1]) for a in m.group(2).split(", "))
lines = {}
for cub in os.listdir(tmp):
    if not cub.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    sec = re.search(r"\.text\.(_ZN2sn\d+%sI%sE\w*):\n(.*?)(?=\n//-{10,} \.text|\Z)" % (fn, enc), txt, re.S)
    if not sec:
        continue
    cur = None
    stack = []
    for ln in sec.group(2).splitlines():
        mm = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if mm:
            cur = (os.path.basename(mm.group(1)), int(mm.group(2)), "inlined" in mm.group(3))
            continue
        mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if mm:
            lines[int(mm.group(1), 16)] = cur
    break
src = {}


def text(f, l):
    if f not in src:
        for root, _, files in os.walk(os.path.dirname(os.path.abspath(lib))):
            if f in files:
                src[f] = open(os.path.join(root, f)).read().splitlines()
                break
        else:
            src[f] = []
    return src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""


print(f"# {kernel}")
total_s = sum(s for _, _, s, _ in sass)
print(f"# stall samples total {total_s}")
print("| addr | first tries | retries | samples near | nearest non-wrapper source line |\n|---|---|---|---|---|")
idx = [i for i, (a, n, s, t) in enumerate(sass) if "PHASECHK" in t and n > 0]
k = 0
while k < len(idx):
    i = idx[k]
    first = sass[i][1]
    retries = 0
    if k + 1 < len(idx) and sass[idx[k + 1]][0] - sass[i][0] < 0x100:
        retries = sass[idx[k + 1]][1]
        k += 1
    near = sum(s for a, n, s, t in sass[max(0, i - 2):i + 14])
    # walk back to the closest instruction whose line is in the kernel file (not the PTX wrapper header)
    j = i
    loc = None
    while j > 0:
        l = lines.get(sass[j][0] - base)
        if l and l[0].endswith(".cu"):
            loc = l
            break
        j -= 1
    print(f"| {sass[i][0]-base:#x} | {first} | {retries} | {near} | {loc[0]}:{loc[1]} `{text(loc[0], loc[1])}` |" if loc else
          f"| {sass[i][0]-base:#x} | {first} | {retries} | {near} | ? |")
    k += 1
