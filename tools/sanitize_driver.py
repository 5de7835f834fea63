"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): golden clusters, a synthetic batch with
dense clusters (exercises every workspace tier), a merge batch and the batched wfa_ed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from aardvark_b200 import synth
from aardvark_b200.batch import RegionBatch
from aardvark_b200.lib import Solver
from aardvark_b200.types import CompareConfig, CompareRegion, MergeConfig
from golden_cases import COMPARE_CASES, MOCK_CHR1

s = Solver(0)
s.set_reference([MOCK_CHR1], ["mock_chr1"])
regions = [CompareRegion(i, r.coordinates, r.truth_variants, r.truth_zygosity, r.query_variants, r.query_zygosity)
           for i, (_, r, _) in enumerate(COMPARE_CASES)]
out = s.compare_batch(RegionBatch.from_compare_regions(regions, s.contig_index), CompareConfig(enable_sequences=True))
print("golden", out.status[:8])
p = synth.SynthParams(n_variants=600, dense_frac=0.8, dense_mean=6.0, het_frac=0.9, phased_frac=0.3, p_repr=0.05, p_gt_err=0.05, p_fn=0.05, p_fp=0.05)
ref, batch = synth.workload_compare(20_000, p, seed=11)
s.set_reference([ref])
out = s.compare_batch(batch, CompareConfig(enable_sequences=False))
print("dense", batch.n_regions, int(out.solved_blocks[0]), s.last_tier_overflow())
ref, mb = synth.workload_merge(30_000, 120, n_sets=3, seed=38)
s.set_reference([ref])
print("merge", s.merge_batch(mb, MergeConfig(majority_voting_enabled=True)).classification[:10])
# SV-sized events: 2 MB global tier -> cooperative tier (named barriers, shared-memory wavefronts)
ps = synth.SynthParams(n_variants=12, sv_events=3, sv_min=600, sv_max=1500, flank=1000)
ref, sv = synth.workload_compare(60_000, ps, seed=44)
s.set_reference([ref])
out = s.compare_batch(sv, CompareConfig(enable_sequences=False))
print("sv", sv.n_regions, int(out.solved_blocks[0]), s.last_tier_overflow())
# device region builder
from aardvark_b200.batch import CallSets
ref, inputs = synth.callsets_compare(50_000, synth.SynthParams(n_variants=150), seed=7)
s.set_reference([ref])
print("builder", s.build_regions(CallSets(inputs), 0, 50).n_regions)
print("wfa", s.wfa_ed_batch([(b"ACGTACGTACGT", b"ACTACGCACGGGT"), (b"A" * 300, b"A" * 150 + b"C" + b"A" * 149)]))
s.close()
