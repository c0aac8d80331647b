"""Read an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv) and print the hottest SASS lines with their
dominant stall reasons.  usage: ncu_hot.py file.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = rows[1]
si = h.index("# Samples")
stall_cols = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
data = []
tot = 0
for r in rows[2:]:
    if len(r) <= si:
        continue
    try:
        n = int(r[si])
    except ValueError:
        continue
    tot += n
    data.append((n, r))
print("total samples", tot)
idx = {id(r): i for i, (n, r) in enumerate(data)}
for n, r in sorted(data, key=lambda t: -t[0])[:top]:
    st = sorted(((int(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:3]
    print(f"{n:7d} {100.0*n/tot:5.1f}%  #{idx[id(r)]:5d} {r[1].strip()[:70]:70s} {[(k[6:], v) for v, k in st if v]}")
