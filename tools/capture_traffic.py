"""profiles/rNN_traffic.json from an ncu --set full capture of one forward (tools/profile_fwd.py <batch>): DRAM bytes
(read + write) of the tcgen05 conv launches, stamped with the digest of the kernel sources of the build that was
captured -- bench.py refuses the figure when the sources have changed since.
usage: capture_traffic.py <report.ncu-rep> <batch> <out.json> [source note]"""
import csv
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import supernet_b200 as S

rep, batch, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
# a report, or the csv of its raw page (`ncu -i rep --page raw --csv`, made on the GPU box when the report is too big)
raw = open(rep).read() if rep.endswith(".csv") else \
    subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
ik, ir, iw = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
conv = total = 0.0
n = 0
for r in rows[2:]:
    b = float(r[ir].replace(",", "")) * scale[u[ir]] + float(r[iw].replace(",", "")) * scale[u[iw]]
    total += b
    if "conv_moments_halo_kernel" in r[ik]:
        conv += b
        n += 1
json.dump({"batch": batch, "tcgen05_conv_launches": n, "dram_bytes_per_step": conv,
           "all_kernels_dram_bytes_per_step": total, "build_digest": S.build._digest(),
           "source": sys.argv[4] if len(sys.argv) > 4 else f"ncu --set full --clock-control none of tools/profile_fwd.py {batch}: "
           "dram__bytes_read.sum + dram__bytes_write.sum over the conv_moments_halo_kernel launches of one forward"},
          open(out, "w"), indent=1)
print(open(out).read())
