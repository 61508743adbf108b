"""Per-kernel table from ONE `ncu --set full` report over tools/kernel_zoo.py --once (the second launch of each kernel is the
measured one; the first is the warm-up).  usage: python tools/ncu_kernels_summary.py <report.ncu-rep> <out.csv>"""
import csv
import io
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
cols = [h for h in WANT if h in hdr]


def num(d, h):
    i = hdr.index(h)
    try:
        v = float(d[i].replace(",", ""))
    except ValueError:
        return d[i]
    u = units[i]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "msecond": 1e3, "second": 1e6, "nsecond": 1e-3}.get(u, 1)


seen = {}
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launch_of_that_kernel"] + cols + ["dram_GBps_in_launch"])
    for d in data:
        name = d[hdr.index("Kernel Name")]
        short = re.sub(r"^void\s+", "", name)
        short = re.sub(r"\(.*$", "", short)
        seen[short] = seen.get(short, 0) + 1
        vals = [num(d, h) for h in cols]
        t_us = num(d, "gpu__time_duration.sum")
        by = num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum")
        w.writerow([short[:110], seen[short]] + [round(v, 3) if isinstance(v, float) else v for v in vals] + [round(by / t_us / 1e3, 1) if t_us else ""])
print(open(out).read()[:6000])
