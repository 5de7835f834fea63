import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aardvark_b200 import synth
from aardvark_b200.lib import Solver
from aardvark_b200.types import CompareConfig
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
ref, b = synth.workload_chr20(scale=scale, seed=20)
s = Solver(0); s.set_reference([ref])
cfg = CompareConfig(enable_sequences=False)
s.upload(b)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    s.run_resident(cfg)
print(s.last_timings_ms(), s.last_tier_ms())
