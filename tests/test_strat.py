"""Stratification containment: the oracle's restatement of Stratifications::containments against the
reference's own golden (src/parsing/stratifications.rs:219-244 on test_data/example_stratification), and the
var_coordinates() quirk (compare_region.rs:63-74: the end comes from the LAST variant of each list)."""
import numpy as np

import oracle_py as orc
from aardvark_b200 import abi
from aardvark_b200.batch import RegionBatch, StratIntervals, masks_to_membership
from aardvark_b200.types import CompareRegion, Coordinates, PhasedZygosity as Z, Variant

# test_data/example_stratification/{example1.bed, example2.bed} (contigs mock = 0, mock2 = 1), labels in strat.tsv order
EXAMPLE1 = [(0, 10, 20), (0, 25, 30), (1, 15, 25)]
EXAMPLE2 = [(0, 15, 25), (0, 26, 30), (1, 10, 25), (1, 30, 35)]
STRAT = StratIntervals([EXAMPLE1, EXAMPLE2], n_contigs=2)


def _query(chrom, first, last):
    """containments(chrom, first, last) through a region whose var_coordinates() are [first, last + 1)"""
    v = Variant(0, abi.VT_DELETION if last > first else abi.VT_SNV, first, b"A" * (last - first + 1), b"C")
    r = CompareRegion(0, Coordinates(chrom, 0, 100), [v], [Z.HomozygousAlternate], [], [])
    b = RegionBatch.from_compare_regions([r], {"mock": 0, "mock2": 1})
    m = int(orc.containments(b, STRAT)[0])
    return [s for s in range(2) if (m >> s) & 1]


def test_example_stratification_golden():   # stratifications.rs:232-243
    assert _query("mock", 9, 9) == []
    assert _query("mock", 10, 10) == [0]
    assert _query("mock", 14, 14) == [0]
    assert _query("mock", 15, 15) == [0, 1]
    assert _query("mock", 20, 20) == [1]
    assert _query("mock", 25, 25) == [0]
    assert _query("mock", 10, 19) == [0]
    assert _query("mock", 15, 24) == [1]
    assert _query("mock2", 15, 24) == [0, 1]
    assert _query("mock2", 30, 34) == [1]
    assert _query("mock2", 9, 9) == []


def test_var_coordinates_uses_the_last_variant_not_the_furthest():   # compare_region.rs:69-71
    # truth: a 9-base deletion at 10 (reaches 19) then an SNV at 12 -> end = 13 from the LAST variant, so [10, 12] is queried
    # and lies inside example1's [10, 19]; the furthest-reaching end (19) would still be inside, so move it past: deletion to 24
    t = [Variant(0, abi.VT_DELETION, 10, b"A" * 15, b"A"), Variant(0, abi.VT_SNV, 12, b"A", b"C")]
    r = CompareRegion(0, Coordinates("mock", 0, 100), t, [Z.HomozygousAlternate, Z.HomozygousAlternate], [], [])
    b = RegionBatch.from_compare_regions([r], {"mock": 0, "mock2": 1})
    assert int(orc.containments(b, STRAT)[0]) == 0b01      # [10, 12] inside example1 [10, 19]; with max() it would be [10, 24]: nowhere
    # the query list's last variant decides too
    q = [Variant(0, abi.VT_SNV, 27, b"A", b"C")]
    r2 = CompareRegion(1, Coordinates("mock", 0, 100), [t[1]], [Z.HomozygousAlternate], q, [Z.HomozygousAlternate])
    b2 = RegionBatch.from_compare_regions([r2], {"mock": 0, "mock2": 1})
    assert int(orc.containments(b2, STRAT)[0]) == 0        # [12, 27] spans both example1 intervals


def test_masks_to_membership():
    off, idx = masks_to_membership(np.array([0b00, 0b11, 0b10], dtype=np.uint64))
    assert off.tolist() == [0, 0, 2, 3] and idx.tolist() == [0, 1, 1]
