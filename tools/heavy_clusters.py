"""Per-cluster device time of the heaviest chr20 clusters (each run as a 1-cluster batch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aardvark_b200 import synth
from aardvark_b200.lib import Solver
from aardvark_b200.types import CompareConfig
ref, b = synth.workload_chr20(1.0, 20)
s = Solver(0); s.set_reference([ref])
cfg = CompareConfig(enable_sequences=False)
ids = [int(x) for x in sys.argv[1:]] or [56342, 38049, 85621, 80623, 97260, 70952, 52297, 110117, 61541, 48159, 8190, 63338, 6562, 57654]
N = (b.var_off[2::2] - b.var_off[0:-2:2]).astype(int)
for r in ids:
    one = b.slice_regions(r, r + 1)
    s.upload(one)
    for _ in range(3): s.run_resident(cfg)
    t = s.last_timings_ms(); w = s.last_work()
    print(r, "N", N[r], "search_ms %.3f" % t["search"], "tiers", [round(x, 3) for x in s.last_tier_ms()], "overflow", s.last_tier_overflow(),
          "spops", w["search_pops"], "xpops", w["exact_pops"], "cells", w["cells"], "aligns", w["alignments"])
