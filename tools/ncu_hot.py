"""Hot SASS lines of one kernel from `ncu -i X.ncu-rep --page source --csv [--print-source sass]` on stdin:
prints the header names once and the lines with the most warp-stall samples."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hi = [i for i, r in enumerate(rows) if any("Sampl" in c for c in r)]
if not hi:
    print("no sampling column found; first rows:", rows[:3]); sys.exit(0)
h = rows[hi[0]]
print("columns:", h)
ci = [i for i, c in enumerate(h) if "Sampl" in c and "Not" not in c][0]
si = [i for i, c in enumerate(h) if c.strip() in ("Source", "SASS", "Instruction")]
si = si[0] if si else 1
body = [r for r in rows[hi[0] + 1:] if len(r) > ci and r[ci].replace(".", "").isdigit()]
tot = sum(float(r[ci]) for r in body)
print("total samples", tot, "lines", len(body))
idx = sorted(range(len(body)), key=lambda i: -float(body[i][ci]))[: int(sys.argv[1]) if len(sys.argv) > 1 else 60]
for i in sorted(idx):
    r = body[i]
    print(f"{i:5d} {float(r[ci]) / tot * 100:6.2f}%  {r[si][:110]}")
