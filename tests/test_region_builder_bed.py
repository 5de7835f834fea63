"""Region builder over BED intervals and several contigs (region_generation.rs:276-479).

The reference has no unit tests for this function (its test module is a TODO, region_generation.rs:815-822), so parity is
UNPINNED by reference vectors: the oracle restatement (orc_build_regions_bed, a literal transcription of the iterator's
deque loop) is checked here against expectations derived by hand from the reference's rules, and the device builder
(avk_build_regions_bed: one radix sort, a binary search per variant, one segmented max-scan) against the oracle on the same
hand-made corner cases and on random multi-contig call sets (tests/test_gpu_parity.py)."""
import numpy as np

import oracle_py as orc
from aardvark_b200 import abi
from aardvark_b200.batch import BedIntervals, CallSets

SNV = abi.VT_SNV
HET = abi.ZYG_UNPHASED_HET


def rec(pos, ref_len=1, alt=b"G"):
    return (pos, b"A" * ref_len, alt, HET, SNV if ref_len == 1 else abi.VT_DELETION, max(ref_len, len(alt)))


def corner_case_callsets():
    """Two inputs on contig 0 (length 1000) and one variant on contig 1 (length 500); flank 10.
    BED: contig 0: [100, 200), [200, 240), [300, 400); contig 1: [0, 50)."""
    truth = [rec(50), rec(100), rec(105, 3), rec(150), rec(198, 5), rec(199), rec(201), rec(250), rec(300), rec(399, 2), rec(450)]
    query = [rec(100), rec(305), rec(20)]          # the last one on contig 1
    cs = CallSets([truth, query], contigs=[[0] * len(truth), [0, 0, 1]])
    bed = BedIntervals([[(100, 200), (200, 240), (300, 400)], [(0, 50)]])
    return cs, bed, [1000, 500]


def test_oracle_bed_builder_hand_derived():
    cs, bed, lens = corner_case_callsets()
    b = orc.build_regions_bed(cs, lens, 10, bed, first_region_id=7)
    # clusters in order: contig 0 interval [100,200): {100 (t, q), 105} window [90, 118); {150} [140, 161); {199} [189, 210)
    #   (198+5 overlaps the interval end: dropped; 50 is before the first interval: dropped)
    # interval [200,240): {201} [191, 212) -- a fresh window per interval, although 201 < 210
    # interval [300,400): 250 is before it: dropped; {300, 305 (q)} [290, 316); 399+2 overlaps: dropped; 450 is after the last interval
    # contig 1 interval [0,50): {20 (q)} window [10, 31)
    assert b.n_regions == 6
    assert list(b.region_id) == [7, 8, 9, 10, 11, 12]                      # one running counter across intervals and contigs
    assert list(b.contig) == [0, 0, 0, 0, 0, 1]
    assert list(b.start) == [90, 140, 189, 191, 290, 10]
    assert list(b.end) == [118, 161, 210, 212, 316, 31]
    vo = b.var_off.astype(int)
    per_region = [[list(b.position[vo[2 * r + k]:vo[2 * r + k + 1]]) for k in range(2)] for r in range(6)]
    assert per_region == [[[100, 105], [100]], [[150], []], [[199], []], [[201], []], [[300], [305]], [[], [20]]]


def test_oracle_bed_builder_window_clipping_and_no_bed():
    """window_start saturates at 0 and is NOT clipped to the interval; window_end is clipped to the contig length only."""
    truth = [rec(3), rec(996, 4)]
    cs = CallSets([truth, []], contigs=[[0, 0], []])
    b = orc.build_regions_bed(cs, [1000], 10, BedIntervals([[(2, 1000)]]))
    assert list(b.start) == [0, 986] and list(b.end) == [14, 1000]
    # without a BED every contig is one interval: same result as the single-contig builder
    b2 = orc.build_regions_bed(cs, [1000], 10, None)
    b1 = orc.build_regions(CallSets([truth, []]), 1000, 10)
    for f in ("start", "end", "var_off", "position", "allele_off"):
        assert np.array_equal(getattr(b1, f), getattr(b2, f)), f
    # a variant reaching past the contig end is dropped (:551)
    cs3 = CallSets([[rec(998, 5)], []], contigs=[[0], []])
    assert orc.build_regions_bed(cs3, [1000], 10, None).n_regions == 0
