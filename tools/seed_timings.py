"""Device time of one chr20-scale pass per seed (bench.py uses seed 20 + rank): shows how much one dense cluster moves a pass."""
import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aardvark_b200 import synth
from aardvark_b200.lib import Solver
from aardvark_b200.types import CompareConfig
s = Solver(0)
cfg = CompareConfig(enable_sequences=False)
for seed in ([int(x) for x in sys.argv[1:]] or range(20, 28)):
    ref, b = synth.workload_chr20(1.0, seed)
    s.set_reference([ref]); s.upload(b)
    for _ in range(4): s.run_resident(cfg)
    t = s.last_timings_ms()
    print(seed, b.n_regions, "total %.3f search %.3f" % (t["total"], t["search"]), "tiers", [round(x,2) for x in s.last_tier_ms()], s.last_tier_overflow())
