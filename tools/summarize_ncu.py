"""Summarise an ncu report (and a launch-list csv) into the tracked profiles/ directory.
usage: summarize_ncu.py <report.ncu-rep> <launches.csv> <out.md> [layer names csv]"""
import csv
import subprocess
import sys

rep, launches, out = sys.argv[1:4]
names = sys.argv[4].split(",") if len(sys.argv) > 4 else None
# a report, or the csv of its raw page (`ncu -i rep --page raw --csv`, made on the GPU box when the report is too big)
raw = open(rep).read() if rep.endswith(".csv") else \
    subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
cols = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("gpu__time_duration.sum", "us"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "umma_pipe%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%"),
        ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%")]
idx = [(h.index(c), n) for c, n in cols if c in h]
units = rows[1]
with open(out, "w") as f:
    f.write(f"# ncu summary of {rep.split('/')[-1]} (one forward, `--set full --clock-control none`; cold-cache, "
            "serialised launches: compare shares, not absolutes)\n\n")
    f.write("| # | layer | " + " | ".join(n for _, n in idx) + " |\n|" + "---|" * (len(idx) + 2) + "\n")
    tot = 0.0
    for k, r in enumerate(rows[2:]):
        vals = []
        for i, n in idx:
            v = r[i]
            if n == "kernel":
                v = v.split("(")[0].replace("void sn::", "").replace("sn::", "")[:44]
            elif n in ("dram_rd_MB", "dram_wr_MB"):
                u = units[i]
                x = float(v.replace(",", ""))
                x = x / 1e6 if u == "byte" else (x / 1e3 if u == "Kbyte" else (x * 1e3 if u == "Gbyte" else x))
                v = f"{x:.1f}"
            elif n == "us":
                x = float(v.replace(",", ""))
                x = x / 1e3 if units[i] == "ns" else x
                tot += x
                v = f"{x:.1f}"
            else:
                try:
                    v = f"{float(v.replace(',', '')):.1f}" if "." in v else v
                except ValueError:
                    pass
            vals.append(v)
        lname = names[k] if names and k < len(names) else ""
        f.write(f"| {k} | {lname} | " + " | ".join(vals) + " |\n")
    f.write(f"\nTotal kernel time in this capture: {tot:.1f} us.\n")
    # launch list
    f.write("\n## launch list (`--metrics gpu__time_duration.sum`)\n\n| # | layer | kernel | us |\n|---|---|---|---|\n")
    lr = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
    if not lr:          # metric-only capture: the table above already carries the per-launch times
        print("wrote", out)
        sys.exit(0)
    lh = lr[0]
    ik, iv, iu = lh.index("Kernel Name"), lh.index("Metric Value"), lh.index("Metric Unit")
    t2 = 0.0
    body = lr[1:]
    for k, r in enumerate(body):
        x = float(r[iv].replace(",", ""))
        x = x / 1e3 if r[iu] in ("ns", "nsecond") else x
        t2 += x
        lname = names[k] if names and k < len(names) else ""
        f.write(f"| {k} | {lname} | {r[ik].split('(')[0].replace('void sn::', '').replace('sn::', '')[:44]} | {x:.1f} |\n")
    f.write(f"\nTotal: {t2:.1f} us; tcgen05 conv share: "
            f"{sum(float(r[iv].replace(',', '')) for r in body if 'halo' in r[ik] or 'conv_moments_tc' in r[ik]) / sum(float(r[iv].replace(',', '')) for r in body):.3f}\n")
print("wrote", out)
