"""Top sampled SASS instructions of the first kernel in an ncu report: python profiles/hotspots.py rep [N]"""
import csv, io, subprocess, sys
rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
s = starts[0]; e = starts[1] if len(starts) > 1 else len(rows)
hdr = rows[s]
si, sm, ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [r for r in rows[s + 1:e] if r and r[0].startswith("0x")]
tot = sum(int(r[sm]) for r in body)
print("total samples", tot, "instructions", len(body))
idx = sorted(range(len(body)), key=lambda i: -int(body[i][sm]))[:n]
for i in sorted(idx):
    r = body[i]
    print(f"{i:5d} {int(r[sm]):7d} {100.0 * int(r[sm]) / tot:5.1f}%  exec {int(r[ex]):10d}  {r[si][:90]}")
