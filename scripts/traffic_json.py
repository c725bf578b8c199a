"""Turns an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`
log of traverse_kernel launches into profiles/<name>.csv + .json (bench.py reads the JSON for
roofline.traffic).  usage: python scripts/traffic_json.py ncu_log.csv workload out_prefix"""
import csv
import json
import sys

src, workload, prefix = sys.argv[1:4]
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.DictReader(lines))
per = {}
for r in rows:
    if "traverse" not in r["Kernel Name"]:
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    per.setdefault(int(r["ID"]), {})[r["Metric Name"]] = v * scale
ids = sorted(per)
with open(prefix + ".csv", "w") as f:
    f.write("launch,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum_ns\n")
    for k, i in enumerate(ids):
        p = per[i]
        f.write(f"{k},{p['dram__bytes_read.sum']:.0f},{p['dram__bytes_write.sum']:.0f},{p['gpu__time_duration.sum']:.0f}\n")
total = sum(per[i]["dram__bytes_read.sum"] + per[i]["dram__bytes_write.sum"] for i in ids)
json.dump({"workload": workload, "kernel": "traverse_kernel",
           "command": "PT_LANES=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:traverse "
                      "--clock-control none python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary",
           "launches": len(ids), "dram_bytes_total": total, "dram_bytes_per_launch": total / max(1, len(ids)),
           "ncu_time_ns_total": sum(per[i]["gpu__time_duration.sum"] for i in ids), "source": prefix + ".csv"},
          open(prefix + ".json", "w"), indent=1)
print(prefix + ".json", len(ids), "launches", total / max(1, len(ids)) / 1e6, "MB per launch")
