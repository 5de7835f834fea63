"""DRAM traffic of one pass from an ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum):
python tools/traffic_summary.py FILE.csv PASSES CONFIG SCALE > profiles/r2_pipeline_traffic.json"""
import collections
import csv
import json
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
passes, config, scale = float(sys.argv[2]), sys.argv[3], float(sys.argv[4])
hdr = rows[0]
ki, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi or not r[mi].startswith("dram__bytes"):
        continue
    v = float(r[vi].replace(",", "")) * mult.get(r[ui], 1.0)
    name = r[ki].split("(")[0][:60]
    if name.startswith("void at::") or "elementwise" in name:
        continue                                   # bench.py's L2 flush, not the library
    per[name] = per.get(name, 0.0) + v
total = sum(per.values())
print(json.dumps({"config": config, "scale": scale, "dram_bytes_per_step": total / passes,
                  "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over every library kernel of the bench command, per pass "
                            f"({passes:g} passes in the capture; profiled offline, cold caches)",
                  "per_kernel_bytes_per_step": {k: v / passes for k, v in sorted(per.items(), key=lambda x: -x[1])}}, indent=1))
