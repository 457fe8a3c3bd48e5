"""Attribute an ncu report's per-SASS-instruction counters to CUDA source lines.
ncu's csv export of the CUDA source view carries no metrics, so: the SASS view of the report (address, instructions
executed, stall samples) is joined with `nvdisasm -g` of the shipped cubin (address -> file:line).
usage: ncu_by_line.py <report.ncu-rep> <library.so> [top N]"""
import csv
import collections
import os
import re
import subprocess
import sys
import tempfile

rep, lib = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
# a report, or the csv of its source page (`ncu -i rep --page source --csv`, made on the GPU box)
raw = open(rep).read() if rep.endswith(".csv") else \
    subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
kernel = rows[0][1]
h = rows[1]
ia, ii, isamp, isrc = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
sass = []
for r in rows[2:]:
    if len(r) < len(h):
        continue
    sass.append((int(r[ia], 16), int(r[ii] or 0), int(r[isamp] or 0), r[isrc]))
base = sass[0][0]
# mangled name of the kernel: match on template arguments
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
m = re.match(r"void sn::(\w+)<(.*)>\(", kernel)
fn, targs = m.group(1), [a.split(")")[-1] for a in m.group(2).split(", ")]
enc = "".join(("Lb%sE" % a) if a in "01" and "bool" in t else ("Li%sE" % a)
              for a, t in zip(targs, m.group(2).split(", ")))
lines = {}
for cub in os.listdir(tmp):
    if not cub.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    sec = re.search(r"\.text\.(_ZN2sn\d+%sI%sE\w*):\n(.*?)(?=\n//-{10,} \.text|\Z)" % (fn, enc), txt, re.S)
    if not sec:
        continue
    cur = None
    for ln in sec.group(2).splitlines():
        mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if mm:
            cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if mm:
            lines[int(mm.group(1), 16)] = cur
    break
by = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for a, n, s, _ in sass:
    key = lines.get(a - base)
    by[key][0] += n
    by[key][1] += s
    tot_i += n
    tot_s += s
src_cache = {}


def text(key):
    if key is None:
        return "?"
    f, l = key
    if f not in src_cache:
        for root, _, files in os.walk(os.path.dirname(os.path.abspath(lib))):
            if f in files:
                src_cache[f] = open(os.path.join(root, f)).read().splitlines()
                break
        else:
            src_cache[f] = []
    s = src_cache[f]
    return s[l - 1].strip()[:100] if 0 < l <= len(s) else ""


print(f"# {kernel}\n# warp instructions executed {tot_i}, stall samples {tot_s}")
print("| inst % | samples % | line | source |\n|---|---|---|---|")
for key, (n, s) in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    loc = f"{key[0]}:{key[1]}" if key else "?"
    print(f"| {100.0 * n / max(tot_i, 1):.1f} | {100.0 * s / max(tot_s, 1):.1f} | {loc} | `{text(key)}` |")
