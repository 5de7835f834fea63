"""Small SV-sized batch (same shape as tests' large-SV case) for profiling the cooperative tier."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aardvark_b200 import synth
from aardvark_b200.lib import Solver
from aardvark_b200.types import CompareConfig
p = synth.SynthParams(n_variants=int(sys.argv[1]) if len(sys.argv) > 1 else 40, sv_events=int(sys.argv[2]) if len(sys.argv) > 2 else 12,
                      sv_min=1500, sv_max=5000, flank=1000)
ref, b = synth.workload_compare(300_000, p, seed=44)
s = Solver(0); s.set_reference([ref])
cfg = CompareConfig(enable_sequences=False)
s.upload(b)
for i in range(2):
    t0 = time.time(); s.run_resident(cfg); dt = time.time() - t0
    t = s.last_timings_ms(); w = s.last_work()
    print(f"run {i}: clusters {b.n_regions} wall {dt:.3f}s search {t['search']:.1f} ms tiers", [round(x, 2) for x in s.last_tier_ms()],
          "overflow", s.last_tier_overflow(), f"cells {w['cells']:.4g} spops {w['search_pops']} xpops {w['exact_pops']} aligns {w['alignments']}")
