"""Summarise an `ncu --set full` report of the fused step kernel into profiles/: a metric CSV per launch, the per-phase
instruction / stall-sample split of the SASS (segments between barriers) and profiles/ncu_traffic.json (DRAM bytes per
launch, read by bench.py for roofline.traffic).   usage: python tools/ncu_summary.py <report.ncu-rep> <tag>"""
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles")   # on the GPU box: gpurun_out/ (the only directory that travels back)


def page(name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


rows = page("raw")
hdr, units, data = rows[0], rows[1], rows[2:]
keep = [h for h in hdr if any(h.startswith(p) for p in (
    "Kernel Name", "gpu__time_duration", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput", "sm__throughput",
    "lts__throughput", "l1tex__throughput", "launch__", "smsp__inst_executed.sum", "sm__warps_active", "smsp__issue_active",
    "smsp__average_warps_issue_stalled", "lts__t_sector_hit_rate", "sm__inst_executed_pipe", "l1tex__data_bank_conflicts"))]
with open(os.path.join(OUT, f"{tag}_ncu_full_summary.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
    for h in keep:
        i = hdr.index(h)
        w.writerow([h, units[i]] + [d[i] for d in data])


def val(d, name):
    i = hdr.index(name)
    v, u = float(d[i]), units[i]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


traffic = {}
for d in data:
    kname = re.sub(r"^void\s+|<.*$|\(.*$", "", d[hdr.index("Kernel Name")]).split("::")[-1]
    traffic[kname] = {"dram_bytes_per_launch": val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum"),
                      "dram_read": val(d, "dram__bytes_read.sum"), "dram_write": val(d, "dram__bytes_write.sum"),
                      "gpu_time_us_under_ncu": float(d[hdr.index("gpu__time_duration.sum")]), "report": os.path.basename(rep), "tag": tag}
json.dump(traffic, open(os.path.join(OUT, "ncu_traffic.json"), "w"), indent=1)

# per-phase split of the SASS: segments delimited by barriers
src = page("source")
h2 = src[1]
body = [r for r in src[2:] if len(r) > 10 and r[0] != "Address"]
si, ci, ii = h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
tot_s = sum(int(r[ci]) for r in body) or 1
tot_i = sum(int(r[ii]) for r in body) or 1
lines = [f"# {tag}: SASS segments between barriers of the first captured launch; samples = warp-state samples (time share),",
         f"# instr = warp-level instructions executed.  total samples {tot_s}, total instructions {tot_i}, {len(body)} SASS lines",
         "end_sass_line,samples,samples_pct,instructions,instructions_pct,delimiter"]
seg_s = seg_i = 0
for n, r in enumerate(body):
    seg_s += int(r[ci]); seg_i += int(r[ii])
    if re.search(r"BAR|UCGABAR|EXIT", r[si]):
        if seg_s > tot_s * 0.003 or seg_i > tot_i * 0.003:
            lines.append(f"{n},{seg_s},{100 * seg_s / tot_s:.1f},{seg_i},{100 * seg_i / tot_i:.1f},{r[si].strip()[:40]}")
        seg_s = seg_i = 0
open(os.path.join(OUT, f"{tag}_ncu_phase_split.csv"), "w").write("\n".join(lines) + "\n")
print(json.dumps(traffic, indent=1))
print("\n".join(lines[:40]))
