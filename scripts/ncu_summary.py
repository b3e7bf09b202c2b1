"""Markdown table of the headline counters of every launch in an ncu report.
usage: python scripts/ncu_summary.py REPORT.ncu-rep [RAW_CSV_OUT] > table.md"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(raw)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]
idx = [hdr.index(w) for w in want if w in hdr]
print("| " + " | ".join(hdr[i] for i in idx) + " |")
print("|" + "---|" * len(idx))
for r in body:
    cells = []
    for i in idx:
        v = r[i]
        if hdr[i] == "Kernel Name":
            v = v[:60]
        else:
            try:
                v = f"{float(v.replace(',', '')):.6g} {units[i]}".strip()
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
