"""SV / long-indel heavy compare (BASELINE configs[3]) at a given scale: device time, executed work, integer roofline."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aardvark_b200 import synth
from aardvark_b200.batch import CompareOutputs
from aardvark_b200.lib import Solver
from aardvark_b200.types import CompareConfig
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
ref, b = synth.workload_sv(scale, 4)
s = Solver(0); s.set_reference([ref])
cfg = CompareConfig(enable_sequences=False)
s.upload(b)
for i in range(2):
    t0 = time.time(); s.run_resident(cfg); dt = time.time() - t0
    t = s.last_timings_ms(); w = s.last_work()
    ops = 6 * w["cells"] + 4 * ((w["matched_bases"] + 15) // 16)
    print(f"run {i}: clusters {b.n_regions} variants {b.n_variants} wall {dt:.3f}s search {t['search']:.1f} ms total {t['total']:.1f} ms tiers",
          [round(x, 2) for x in s.last_tier_ms()], "overflow", s.last_tier_overflow(),
          f"cells {w['cells']:.4g} Gcells/s {w['cells'] / t['search'] / 1e6:.1f} int_ops {ops:.4g} frac {ops / (t['search'] * 1e-3) / s.int_peak_ops_per_s():.4f}",
          "spops", w["search_pops"], "xpops", w["exact_pops"])
out = s.download(CompareOutputs(b, region_metrics=False))
print("solved", int(out.solved_blocks[0]), "errors", int(out.error_blocks[0]))
