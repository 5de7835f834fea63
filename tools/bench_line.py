"""One-line digest of bench.py output files: python tools/bench_line.py FILE.json ..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        cb = d.get("cpu_baseline", {})
        print(f.split("/")[-1], "ms", round(d["ms_per_step"], 2), "single_pass_ms", round(d.get("single_pass", {}).get("ms_per_step", 0), 2),
              "in_flight", d["config"].get("passes_in_flight"), "e2e_ms", round(d["e2e"]["ms_per_step"], 2), "e2e_single_call_ms", round(d["e2e"].get("single_call_ms", 0), 2),
              {k: round(v, 2) for k, v in d.get("phases_ms", {}).items()}, "cpu_s", cb.get("seconds_per_genome"), "cores", cb.get("cores"),
              "bit_exact", cb.get("matches_gpu_bit_exact"))
    except Exception as e:  # noqa: BLE001
        print(f, "failed:", e)
