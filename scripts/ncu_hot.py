"""Top stall-sample SASS lines of one kernel from an ncu report's source page.
usage: python scripts/ncu_hot.py REPORT.ncu-rep KERNEL_REGEX [launch_skip] [top_n]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
print(rows[hi - 1][1][:150] if hi > 0 else "")
S = hdr.index("# Samples"); SRC = hdr.index("Source"); EX = hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or not r[0].startswith("0x"):
        break
    body.append(r)
tot = sum(int(r[S]) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {}
for r in body:
    for i in stalls:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("stall mix:", {k: round(100 * v / max(tot, 1), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(body)), key=lambda i: -int(body[i][S]))[:top]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[j] or 0), hdr[j][6:]) for j in stalls), reverse=True)[:2]
    print(f"{i:5d} {100 * int(r[S]) / tot:5.1f}%  ex={r[EX]:>9}  {r[SRC].strip()[:90]:90s} {st}")
