"""Prints a markdown table of the metrics that matter from one or more .ncu-rep files (ncu --page raw --csv).
Usage: python profiles/summarize_ncu.py gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...] > profiles/summary.md"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_elapsed.avg", "SM cycles"),
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("<unnamed>::", "").replace("void ", "")
    return name[:70]


seen = {}
for path in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        key = short(r[idx["Kernel Name"]])
        # keep the LONGEST launch of each kernel (the full-size one when a script also launches small ones)
        t = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
        if key not in seen or t > seen[key][0]:
            seen[key] = (t, r, idx, units, path)
print("| kernel (longest captured launch) | " + " | ".join(lbl for _, lbl in METRICS) + " |")
print("|---|" + "---|" * len(METRICS))
for key, (t, r, idx, units, path) in seen.items():
    cells = []
    for m, _ in METRICS:
        if m not in idx:
            cells.append("-")
            continue
        v, u = r[idx[m]], units[idx[m]]
        try:
            f = float(v.replace(",", ""))
            v = f"{f:,.0f}" if abs(f) >= 1000 else f"{f:.2f}".rstrip("0").rstrip(".")
        except ValueError:
            pass
        cells.append(f"{v} {u}".strip())
    print(f"| `{key}` | " + " | ".join(cells) + " |")
