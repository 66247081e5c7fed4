"""Per-CUDA-source-line executed-instruction histogram from
`ncu -i X.ncu-rep --page source --csv --print-source sass,cuda --launch-skip K --launch-count 1`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
cur, hdr, agg, seen_fn = None, None, {}, 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        ie = r.index("Instructions Executed")
        continue
    if hdr is None or len(r) < 10 or r[2] != "-":
        continue
    n = int(r[ie])
    if n:
        key = (cur, int(r[0]))
        agg[key] = (agg.get(key, (0, ""))[0] + n, r[1].strip()[:80])
tot = sum(v[0] for v in agg.values())
print("total warp instructions", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print("%-12s %4d %5.1f%% %s %s" % (k[0], k[1], 100 * v[0] / tot, ("%7.2f/unit" % (v[0] * 32 / units)) if units else "", v[1]))
