"""BASELINE configs[4] (merge x5, majority voting) on one chr20-sized contig: C-ABI wall time vs the CPU oracle, bit-exact check."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import oracle_py as orc
from aardvark_b200 import synth
from aardvark_b200.lib import Solver, merge_cfg
from aardvark_b200.types import MergeConfig
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
ref, b = synth.workload_merge(int(64_444_167 * scale), int(150_000 * scale), n_sets=5, seed=38)
s = Solver(0); s.set_reference([ref])
cfg = MergeConfig(majority_voting_enabled=True)
for i in range(3):
    t0 = time.time(); gpu = s.merge_batch(b, cfg); dt = time.time() - t0
print("tiers ms", [round(x, 2) for x in s.last_tier_ms()], "overflow", s.last_tier_overflow(), "work", s.last_work())
print(f"merge x5: {b.n_regions} clusters, {b.n_variants} variants; GPU (host buffers through the C ABI) {dt * 1e3:.1f} ms -> {b.n_regions / dt / 1e6:.2f} M clusters/s; device", s.last_timings_ms())
t0 = time.time(); cpu = orc.merge_batch(b, [ref], merge_cfg(cfg), n_threads=orc.num_threads()); dc = time.time() - t0
print(f"oracle {dc:.2f} s on {orc.num_threads()} threads -> {b.n_regions / dc / 1e6:.2f} M clusters/s; diff {gpu.diff(cpu)}")
