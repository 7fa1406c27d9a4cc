"""ncu launch list (csv from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`) ->
per-kernel summary JSON read by bench.py for `roofline.traffic`.  usage: summarize_launches.py in.csv out.json"""
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
rows, hdr = [], None
with open(src) as f:
    for r in csv.reader(f):
        if hdr is None:
            if len(r) > 5 and r[0] == "ID":
                hdr = r
            continue
        rows.append(r)
iK, iM, iV, iID = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
iU = hdr.index("Metric Unit")
launch = {}
for r in rows:
    d = launch.setdefault(r[iID], {"name": r[iK]})
    v = float(r[iV].replace(",", ""))
    u = r[iU]
    if r[iM] == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)          # -> ms
    else:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)      # -> bytes
    d[r[iM]] = v
out = {}
for d in launch.values():
    name = re.sub(r"\(.*", "", d["name"])
    o = out.setdefault(name, dict(ms=0.0, launches=0, dram_read_bytes=0.0, dram_write_bytes=0.0))
    o["ms"] += d.get("gpu__time_duration.sum", 0.0)
    o["launches"] += 1
    o["dram_read_bytes"] += d.get("dram__bytes_read.sum", 0.0)
    o["dram_write_bytes"] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(o["ms"] for o in out.values())
for o in out.values():
    o["share"] = o["ms"] / tot
json.dump(dict(sorted(out.items(), key=lambda kv: -kv[1]["ms"])), open(dst, "w"), indent=1)
print(f"{len(launch)} launches, {tot:.2f} ms serialised")
for k, o in list(sorted(out.items(), key=lambda kv: -kv[1]["ms"]))[:12]:
    print(f"{o['ms']:9.3f} ms {100 * o['share']:5.1f}% x{o['launches']:5d}  {(o['dram_read_bytes'] + o['dram_write_bytes']) / 1e9:8.2f} GB  {k[:80]}")
