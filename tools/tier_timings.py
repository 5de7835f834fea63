import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aardvark_b200 import synth, abi
from aardvark_b200.lib import Solver
from aardvark_b200.types import CompareConfig
ref, b = synth.workload_chr20(scale=float(sys.argv[1]) if len(sys.argv)>1 else 0.05, seed=20)
s = Solver(0); s.set_reference([ref])
cfg = CompareConfig(enable_sequences=False)
nt=(b.var_off[1::2]-b.var_off[0:-1:2]).astype(int); nq=(b.var_off[2::2]-b.var_off[1::2]).astype(int)
N = nt+nq
def run(batch, name):
    s.upload(batch)
    for _ in range(3): s.run_resident(cfg)
    t=s.last_timings_ms(); print(name, batch.n_regions, "search %.3f total %.3f"%(t["search"],t["total"]), "tier_ms", [round(x,3) for x in s.last_tier_ms()], "overflow", s.last_tier_overflow())
run(b, "all")
# only N==2 clusters: build via slice of single regions concatenated -> use first such region repeated
idx = np.where(N == 2)[0]
r = int(idx[0])
one = b.slice_regions(r, r+1)
run(one, "single N=2")
for n in (4, 8, 12):
    idx = np.where(N == n)[0]
    if len(idx):
        r = int(idx[0]); run(b.slice_regions(r, r+1), f"single N={n}")
# contiguous slices
run(b.slice_regions(0, 100), "first100")
run(b.slice_regions(0, 3000), "first3000")
