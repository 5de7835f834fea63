"""Print the key fields of a bench.py output file (tolerates extra stdout lines before the JSON line)."""
import json, sys
for path in sys.argv[1:]:
    for l in open(path):
        if l.startswith("{"):
            d = json.loads(l)
            print(path, "n_gpus", d.get("n_gpus"), "value %.4g" % d["value"], d["unit"], "ms_per_step %.3f" % d.get("ms_per_step", float("nan")),
                  "| e2e", d.get("e2e"), "| bit-exact vs CPU:", (d.get("cpu_baseline") or {}).get("matches_gpu_bit_exact"), "| clocks", d.get("clocks"))
        elif l.strip():
            print(path, "extra stdout line:", l.strip()[:80])
