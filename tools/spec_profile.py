"""Per-cluster phase times of k_search_spec on a chr20 pass (AVK_SPEC_PROFILE=1): which phase makes the slowest clusters slow."""
import ctypes as C, os, sys
os.environ["AVK_SPEC_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aardvark_b200 import synth
from aardvark_b200.lib import Solver
from aardvark_b200.types import CompareConfig
ref, b = synth.workload_chr20(1.0, 20)
s = Solver(0); s.set_reference([ref]); s.upload(b)
cfg = CompareConfig(enable_sequences=False)
for _ in range(3): s.run_resident(cfg)
n = 4096
out = np.zeros((n, 8), dtype=np.uint64)
s._lib.avk_spec_profile.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_uint32]
assert s._lib.avk_spec_profile(s._ctx, out.ctypes.data_as(C.POINTER(C.c_uint64)), n) == 0
rows = out[out[:, 1] > 0]
tot = rows[:, 3:7].sum(axis=1)
order = np.argsort(-tot.astype(np.int64))
print("clusters profiled", len(rows), "(the side-stream launch overwrites low slots of the dense launch)")
print("region N pops | load search score metrics+commit (us at 1.965 GHz) | nres")
for i in order[:15]:
    r = rows[i]
    print(int(r[0]), int(r[1]), int(r[2]), "|", " ".join(f"{float(x) / 1965.0:8.1f}" for x in r[3:7]), "|", int(np.int64(r[7])))
print("sum over clusters (ms):", [round(float(rows[:, k].sum()) / 1.965e6, 2) for k in range(3, 7)])
