"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): time per kernel name, share of the total."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
cnt = collections.Counter()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("b2u::", "").replace("void ", "")
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] == "ns" else (v / 1e3 if r[ui] in ("us", "usecond") else v)
    tot[name] = tot.get(name, 0.0) + v
    cnt[name] += 1
total = sum(tot.values())
print(f"total {total:.3f} ms over {sum(cnt.values())} launches (cold-cache, serialised: compare SHARES)")
for k, v in sorted(tot.items(), key=lambda t: -t[1]):
    print(f"{v:9.3f} ms {100*v/total:5.1f}%  x{cnt[k]:4d}  {k}")
