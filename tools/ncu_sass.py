"""Per-SASS-instruction profile of the first kernel in an .ncu-rep, in chunks.
python tools/ncu_sass.py <rep> [chunk] [--full lo hi]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
chunk = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 32
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hdr_i[0]]
end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
ix = {n: i for i, n in enumerate(h)}
data = [r for r in rows[hdr_i[0] + 1:end] if len(r) > ix["Instructions Executed"]]
ie = [float(r[ix["Instructions Executed"]] or 0) for r in data]
sm = [float(r[ix["# Samples"]] or 0) for r in data]
tot, tots = sum(ie), sum(sm)
print("total warp-inst", tot, "samples", tots, "sass lines", len(data))
if "--full" in sys.argv:
    lo, hi = int(sys.argv[sys.argv.index("--full") + 1]), int(sys.argv[sys.argv.index("--full") + 2])
    for k in range(lo, min(hi, len(data))):
        print(f"{k:4d} {ie[k]/tot*100:5.2f}% {sm[k]/tots*100:5.2f}%  {data[k][ix['Source']][:100]}")
    sys.exit()
for i in range(0, len(data), chunk):
    a, b = sum(ie[i:i + chunk]) / tot * 100, sum(sm[i:i + chunk]) / tots * 100
    if a < 0.4 and b < 0.4:
        continue
    ops = collections.Counter()
    for r in data[i:i + chunk]:
        t = r[ix["Source"]].split()
        ops[t[1] if t[0].startswith("@") else t[0]] += 1
    print(f"{i:4d}-{i+chunk-1:4d} inst {a:5.1f}% samp {b:5.1f}%  {dict(ops.most_common(7))}")
