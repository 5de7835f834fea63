"""BASELINE configs[2] (WGS HG002-like: 24 contigs with GRCh38 lengths, ~5 M variants per side) at a length scale:
one multi-contig batch through the C ABI, bit-exact against the CPU oracle, plus the sharding property
(contiguous region bins solved separately == the whole batch)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import oracle_py as orc
from aardvark_b200 import synth
from aardvark_b200.batch import RegionBatch, CompareOutputs
from aardvark_b200.dist import partition_regions
from aardvark_b200.lib import Solver, compare_cfg
from aardvark_b200.types import CompareConfig

GRCH38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422,
          135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167,
          46709983, 50818468, 156040895, 57227415]
f = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
t0 = time.time()
parts = []
for i, L in enumerate(GRCH38):
    Ls = max(20000, int(L * f))
    parts.append(synth.workload_compare(Ls, synth.SynthParams(n_variants=max(1, int(5_000_000 * Ls / 3.1e9))), 38 + i))
contigs = [p[0] for p in parts]
batch = RegionBatch.concat([p[1] for p in parts])
print(f"scale {f}: {len(contigs)} contigs, {sum(c.size for c in contigs) / 1e6:.0f} Mbp, {batch.n_regions} clusters, {batch.n_variants} variants, generated in {time.time() - t0:.0f} s")
s = Solver(0); s.set_reference(contigs)
cfg = CompareConfig(enable_sequences=False)
s.upload(batch)
for _ in range(3):
    s.run_resident(cfg, region_metrics=True)
t = s.last_timings_ms()
print(f"device: total {t['total']:.2f} ms, search {t['search']:.2f} ms -> {batch.n_regions / t['total'] / 1e3:.2f} M clusters/s; tiers",
      [round(x, 2) for x in s.last_tier_ms()], "overflow", s.last_tier_overflow())
gpu = s.download(CompareOutputs(batch))
t0 = time.time()
cpu = orc.compare_batch(batch, contigs, compare_cfg(cfg), n_threads=orc.num_threads())
dt = time.time() - t0
print(f"oracle: {dt:.1f} s on {orc.num_threads()} threads -> {batch.n_regions / dt / 1e6:.2f} M clusters/s; diff vs GPU: {gpu.diff(cpu)}")
# sharding property: 8 contiguous bins solved separately reproduce the whole batch
bins = partition_regions(batch, 8)
ok = True
tot = np.zeros_like(gpu.totals)
for lo, hi in bins:
    o = s.compare_batch(batch.slice_regions(lo, hi), cfg)
    ok &= np.array_equal(o.region_metrics[:hi - lo], gpu.region_metrics[lo:hi]) and np.array_equal(o.status[:hi - lo], gpu.status[lo:hi])
    tot += o.totals
print("8 contiguous bins == whole batch:", bool(ok), "; sum of bin totals == totals:", bool(np.array_equal(tot, gpu.totals)),
      "; solved", int(gpu.solved_blocks[0]), "errors", int(gpu.error_blocks[0]))
