"""Host-side batch plumbing (no GPU): multi-contig concatenation, region slicing and the rank partition, checked through
the CPU oracle -- a contig bin solved on its own must give the rows it has inside the whole batch."""
import numpy as np

import oracle_py as orc
from aardvark_b200 import synth
from aardvark_b200.batch import CallSets, RegionBatch
from aardvark_b200.dist import partition_regions
from aardvark_b200.lib import compare_cfg
from aardvark_b200.types import CompareConfig


def test_concat_slice_and_partition_are_consistent():
    parts = [synth.workload_chr20(scale=sc, seed=sd) for sc, sd in ((0.003, 101), (0.0012, 102), (0.002, 103))]
    contigs = [p[0] for p in parts]
    batch = RegionBatch.concat([p[1] for p in parts])
    assert batch.n_regions == sum(p[1].n_regions for p in parts)
    assert np.array_equal(batch.region_id, np.arange(batch.n_regions, dtype=np.uint64))
    cfg = compare_cfg(CompareConfig(enable_sequences=False))
    whole = orc.compare_batch(batch, contigs, cfg)
    assert int(whole.error_blocks[0]) == 0
    lo = 0
    for ref, b in parts:                                  # each contig alone == its rows in the joined batch
        alone = orc.compare_batch(b, [ref], cfg)
        assert np.array_equal(alone.region_metrics[:b.n_regions], whole.region_metrics[lo:lo + b.n_regions])
        lo += b.n_regions
    bins = partition_regions(batch, 4)                    # contiguous, covering, in order
    assert bins[0][0] == 0 and bins[-1][1] == batch.n_regions and all(a[1] == b[0] for a, b in zip(bins, bins[1:]))
    tot = np.zeros_like(whole.totals)
    for lo, hi in bins:
        part = orc.compare_batch(batch.slice_regions(lo, hi), contigs, cfg)
        assert np.array_equal(part.region_metrics[:hi - lo], whole.region_metrics[lo:hi])
        tot += part.totals
    assert np.array_equal(tot, whole.totals)


def test_callsets_table_round_trips_through_the_builder_oracle():
    ref, inputs = synth.callsets_compare(60_000, synth.SynthParams(n_variants=150), seed=9)
    cs = CallSets(inputs)
    assert cs.n_inputs == 2 and cs.n_variants == sum(len(x) for x in inputs)
    b = orc.build_regions(cs, len(ref), 50)
    assert b.n_variants == cs.n_variants                   # nothing dropped: every record lies inside the contig
    # gap rule: every variant lies inside its window, and the first variant of a cluster is at or past the previous window's end
    vo = b.var_off.astype(np.int64)
    for k in range(b.n_regions):
        pos = b.position[vo[2 * k]:vo[2 * k + 2]].astype(np.int64)
        assert pos.size > 0 and pos.min() >= b.start[k] and (pos + b.a0_len[vo[2 * k]:vo[2 * k + 2]]).max() <= b.end[k]
        if k:
            assert pos.min() >= b.end[k - 1]


def _partition_restated(batch, n_bins):
    """avk_partition_regions in numpy: cuts[k] = first region whose cumulative cost reaches k/n of the total."""
    k = batch.n_inputs
    vo = batch.var_off.astype(np.int64)
    v0, v1 = vo[0:-1:k][:batch.n_regions], vo[k::k]
    win = np.maximum(batch.end.astype(np.int64) - batch.start.astype(np.int64), 0)
    m2 = np.maximum(batch.a0_len, batch.a1_len).astype(np.int64) ** 2
    cs = np.concatenate([[0], np.cumsum(m2)])
    cost = (v1 - v0 + 1) * win + (cs[v1] - cs[v0])
    cum = np.cumsum(cost)
    total = int(cum[-1]) if cum.size else 0
    cuts = [0]
    for j in range(1, n_bins):
        r = int(np.searchsorted(cum * n_bins, total * j, side="left")) if cum.size else 0
        cuts.append(max(min(r, batch.n_regions), cuts[-1]))
    cuts.append(batch.n_regions)
    return list(zip(cuts[:-1], cuts[1:]))


def test_partition_matches_its_definition_small_and_threaded():
    """The library sums the cost proxy per chunk of 16 Ki regions (several host threads from 500 k regions up) and looks a
    cut up inside its chunk: same cuts as the one-pass definition, at sizes on both sides of the chunk and thread limits."""
    _, small = synth.workload_chr20(scale=0.002, seed=5)
    for n_bins in (1, 2, 3, 7, 64):
        assert partition_regions(small, n_bins) == _partition_restated(small, n_bins)
    _, mid = synth.workload_chr20(scale=0.3, seed=6)                 # ~35 k regions: several chunks, one thread
    for n_bins in (2, 5, 8):
        assert partition_regions(mid, n_bins) == _partition_restated(mid, n_bins)
    big = RegionBatch.concat([mid] * 15)                              # > 500 k regions: the threaded path
    assert big.n_regions >= 500000
    for n_bins in (2, 3, 8):
        assert partition_regions(big, n_bins) == _partition_restated(big, n_bins)
