"""Hottest SASS instructions of an ncu report (source page): python tools/ncu_hot.py report.ncu-rep [N]"""
import csv, subprocess, sys
rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
data = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break                                   # first captured launch only
    if len(r) == len(h):
        data.append(r)
col = {k: h.index(k) for k in ("Address", "Source", "# Samples", "stall_long_sb", "Instructions Executed", "stall_wait", "stall_math", "stall_short_sb", "stall_lg", "stall_mio")}
g = lambda r, k: int(float(r[col[k]] or 0))
tot = sum(g(r, "# Samples") for r in data)
print("kernel:", rows[0][1][:100], "| total samples", tot, "| instructions", len(data))
top = sorted(range(len(data)), key=lambda i: -g(data[i], "# Samples"))[:n]
print(f"{'idx':>5} {'samp%':>6} {'long_sb':>7} {'wait':>5} {'math':>5} {'short':>5} {'lg':>4} {'mio':>4} {'exec':>9}  sass")
for i in sorted(top):
    r = data[i]
    print(f"{i:5d} {100*g(r,'# Samples')/tot:6.2f} {g(r,'stall_long_sb'):7d} {g(r,'stall_wait'):5d} {g(r,'stall_math'):5d} {g(r,'stall_short_sb'):5d} {g(r,'stall_lg'):4d} {g(r,'stall_mio'):4d} {g(r,'Instructions Executed'):9d}  {r[col['Source']][:100]}")
