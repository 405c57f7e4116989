"""Top warp-stall-sample SASS lines of an ncu source-page CSV:  ncu -i X.ncu-rep --page source --csv > X.csv; python tools/ncu_top.py X.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]


def f(r, k):
    try:
        return float(r[ix[k]])
    except ValueError:
        return 0.0


tot = sum(f(r, "# Samples") for r in data)
print("rows", len(data), "total samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:n]:
    s = sorted(((k, f(r, k)) for k in stalls), key=lambda kv: -kv[1])[:2]
    print(f'{r[ix["Address"]][-5:]} {f(r, "# Samples"):7.0f} {100 * f(r, "# Samples") / tot:5.1f}% exec={f(r, "Instructions Executed"):8.0f} '
          f'{r[ix["Source"]][:72]:72s} {s}')
