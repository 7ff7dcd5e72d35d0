"""profiles/traffic_r02.json from an extracted ncu raw CSV (tools/ncu_extract.py): DRAM bytes per launch of each kernel function.
    python tools/ncu_traffic.py profiles/ncu_fp32_r02_raw.csv > profiles/traffic_r02.json"""
import csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in rows[2:]:
    m = re.search(r"(k_\w+)", r[ix["Kernel Name"]])
    if not m:
        continue
    k = m.group(1)
    e = out.setdefault(k, {"kernel": k, "launches_captured": 0, "dram": 0.0, "us": 0.0, "inst": 0.0})
    e["launches_captured"] += 1
    for col in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        e["dram"] += float(r[ix[col]]) * scale[units[ix[col]]]
    e["us"] += float(r[ix["gpu__time_duration.sum"]]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}[units[ix["gpu__time_duration.sum"]]]
    e["inst"] += float(r[ix["smsp__inst_executed.sum"]])
res = {}
for k, e in out.items():
    res[k] = {"kernel": k, "launches_captured": e["launches_captured"], "dram_bytes_per_launch": int(e["dram"] / e["launches_captured"]),
              "dram_bytes_per_forward": int(e["dram"]), "ncu_us_per_forward": round(e["us"], 1), "warp_inst_per_forward": int(e["inst"]),
              "source": f"{sys.argv[1]} (ncu --set full, one classifier forward at batch 256, fp32 accuracy mode)"}
print(json.dumps(res, indent=1))
