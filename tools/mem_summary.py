"""Summarise an ncu CSV with gpu__time_duration.sum + dram__bytes_{read,write}.sum per launch: per kernel name the total
time, DRAM traffic and achieved DRAM GB/s; with -v every launch.  usage: mem_summary.py file.csv [-v] [name filter]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
verbose = "-v" in sys.argv
flt = [a for a in sys.argv[2:] if a not in ("-v", "--json")]
hdr = rows[0]
idi, ki, mi, vi, ui = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
L = collections.OrderedDict()
for r in rows[1:]:
    d = L.setdefault(r[idi], {"name": r[ki].split("(")[0].replace("b2u::", "").replace("void ", ""), "grid": r[gi] if gi is not None else ""})
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if r[mi].startswith("gpu__time"):
        d["us"] = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    elif r[mi].startswith("dram__bytes"):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d["rd" if "read" in r[mi] else "wr"] = v * mult
    elif r[mi].startswith("lts__t_bytes"):
        d["l2"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    elif r[mi].startswith("sm__pipe_tensor"):
        d["tc"] = v
agg = collections.OrderedDict()
for d in L.values():
    if flt and not any(f in d["name"] for f in flt):
        continue
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("us", 0); a[2] += d.get("rd", 0); a[3] += d.get("wr", 0)
    if verbose:
        b = d.get("rd", 0) + d.get("wr", 0)
        print(f"{d['name'][:28]:28s} grid {d['grid']:>12s} {d.get('us', 0):9.1f} us  rd {d.get('rd', 0)/1e6:8.1f} MB wr {d.get('wr', 0)/1e6:8.1f} MB  {b / max(d.get('us', 1e-9), 1e-9) / 1e3:7.0f} GB/s")
if "--json" in sys.argv:
    import json
    out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,"
                     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none over the GEMM launches of "
                     "one eager training step (tools/step_eager.py, batch 64); per launch = total / launches"}
    per = {}
    for d in L.values():
        k = d["name"].split("<")[0]
        a = per.setdefault(k, {"launches": 0, "us": 0.0, "dram": 0.0, "l2": 0.0, "tc_w": 0.0})
        a["launches"] += 1; a["us"] += d.get("us", 0); a["dram"] += d.get("rd", 0) + d.get("wr", 0); a["l2"] += d.get("l2", 0)
        a["tc_w"] += d.get("tc", 0) * d.get("us", 0)
    for k, a in per.items():
        out[k] = {"launches": a["launches"], "dram_bytes_per_launch": a["dram"] / a["launches"], "dram_bytes_per_step": a["dram"],
                  "l2_bytes_per_step": a["l2"], "time_ms_cold_serialised": a["us"] / 1e3,
                  "tensor_pipe_active_pct_time_weighted": a["tc_w"] / max(a["us"], 1e-9)}
    print(json.dumps(out, indent=1))
    sys.exit(0)
tot = sum(a[1] for a in agg.values())
print(f"total {tot/1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches")
for k, a in sorted(agg.items(), key=lambda t: -t[1][1]):
    print(f"{a[1]/1e3:9.3f} ms {100*a[1]/tot:5.1f}% x{a[0]:4d}  rd {a[2]/1e9:7.3f} GB wr {a[3]/1e9:7.3f} GB  {(a[2]+a[3])/max(a[1],1e-9)/1e3:7.0f} GB/s  {k}")
