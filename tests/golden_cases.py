"""Golden vectors of the reference's own unit tests / doctests for the hot path, as data.

Each case cites the reference test it restates (paths relative to the reference
repo).  Used twice: tests/test_oracle_golden.py pins the CPU oracle against
them, tests/test_gpu_parity.py runs the same cases through the CUDA C ABI.
"""
from aardvark_b200.types import (CompareRegion, Coordinates, MergeClassification, MergeConfig, MultiRegion,
                                 PhasedZygosity as Z, SummaryGtMetrics, SummaryMetrics, Variant,
                                 VariantMetrics, VariantSource, VariantType)

MOCK_CHR1 = b"ACCGTTACCAGGACTTGACAAACCG"   # src/waffle_solver.rs:806-812, src/merge_solver.rs:234-240

T, Q = VariantSource.Truth, VariantSource.Query
VM = VariantMetrics.new


def _coords(s, e):
    return Coordinates("mock_chr1", s, e)


# (name, CompareRegion, expectations) -- src/waffle_solver.rs:896-1247 and doctests src/lib.rs:9-63
COMPARE_CASES = []


def _case(name, region, **exp):
    COMPARE_CASES.append((name, region, exp))


# src/waffle_solver.rs:897-930 test_solve_compare_region_simple_snv (== doctest src/lib.rs:9-63, waffle_solver.rs:9-62)
_v = [Variant.new_snv(0, 2, b"C", b"G")]
_case("simple_snv", CompareRegion(0, _coords(0, 10), _v, [Z.PhasedHet10], list(_v), [Z.PhasedHet10]),
      total_ed=0, gt=SummaryGtMetrics(1, 0, 1, 0, 0, 0), hap=SummaryMetrics(1, 0, 1, 0),
      basepair=SummaryMetrics(2, 0, 2, 0),
      truth=[VM(T, 1, 1)], query=[VM(Q, 1, 1)],
      seqs=("ACCGTTACCA", "ACGGTTACCA", "ACCGTTACCA", "ACGGTTACCA", "ACCGTTACCA"))

# :934-978 test_solve_compare_region_double_single
_case("double_single",
      CompareRegion(0, _coords(0, 10),
                    [Variant.new_snv(0, 2, b"C", b"G"), Variant.new_snv(0, 3, b"G", b"T")], [Z.PhasedHet10, Z.PhasedHet10],
                    [Variant.new_indel(0, 2, b"CG", b"GT")], [Z.PhasedHet01]),
      total_ed=0, gt=SummaryGtMetrics(2, 0, 1, 0, 0, 0), hap=SummaryMetrics(2, 0, 1, 0),
      basepair=SummaryMetrics(4, 0, 4, 0),
      truth=[VM(T, 1, 1), VM(T, 1, 1)], query=[VM(Q, 1, 1)],
      seqs=("ACCGTTACCA", "ACGTTTACCA", "ACCGTTACCA", "ACGTTTACCA", "ACCGTTACCA"))

# :981-1025 test_solve_compare_region_single_double
_case("single_double",
      CompareRegion(0, _coords(0, 10),
                    [Variant.new_indel(0, 2, b"CG", b"GT")], [Z.PhasedHet01],
                    [Variant.new_snv(0, 2, b"C", b"G"), Variant.new_snv(0, 3, b"G", b"T")], [Z.PhasedHet10, Z.PhasedHet10]),
      total_ed=0, gt=SummaryGtMetrics(1, 0, 2, 0, 0, 0), hap=SummaryMetrics(1, 0, 2, 0),
      basepair=SummaryMetrics(4, 0, 4, 0),
      truth=[VM(T, 1, 1)], query=[VM(Q, 1, 1), VM(Q, 1, 1)],
      seqs=("ACCGTTACCA", "ACCGTTACCA", "ACGTTTACCA", "ACCGTTACCA", "ACGTTTACCA"))

# :1028-1067 test_solve_compare_region_insertion_shift
_case("insertion_shift",
      CompareRegion(0, _coords(0, 10),
                    [Variant.new_insertion(0, 0, b"A", b"AC")], [Z.HomozygousAlternate],
                    [Variant.new_insertion(0, 2, b"C", b"CC")], [Z.HomozygousAlternate]),
      total_ed=0, gt=SummaryGtMetrics(1, 0, 1, 0, 0, 0), hap=SummaryMetrics(2, 0, 2, 0),
      basepair=SummaryMetrics(4, 0, 4, 0),
      truth=[VM(T, 2, 2)], query=[VM(Q, 2, 2)],
      seqs=("ACCGTTACCA", "ACCCGTTACCA", "ACCCGTTACCA", "ACCCGTTACCA", "ACCCGTTACCA"))

# :1070-1103 test_solve_compare_region_simple_fn
_case("simple_fn",
      CompareRegion(0, _coords(0, 10), [Variant.new_snv(0, 2, b"C", b"G")], [Z.PhasedHet10], [], []),
      total_ed=1, gt=SummaryGtMetrics(0, 1, 0, 0, 0, 0), hap=SummaryMetrics(0, 1, 0, 0),
      basepair=SummaryMetrics(0, 2, 0, 0),
      truth=[VM(T, 1, 0)], query=[],
      seqs=("ACCGTTACCA", "ACGGTTACCA", "ACCGTTACCA", "ACCGTTACCA", "ACCGTTACCA"))

# :1105-1156 test_solve_compare_region_complex_001
_case("complex_001",
      CompareRegion(0, _coords(0, 10),
                    [Variant.new_snv(0, 2, b"C", b"G"), Variant.new_snv(0, 4, b"T", b"C"), Variant.new_snv(0, 6, b"A", b"C")],
                    [Z.PhasedHet10, Z.PhasedHet01, Z.PhasedHet10],
                    [Variant.new_snv(0, 2, b"C", b"G"), Variant.new_snv(0, 6, b"A", b"C")],
                    [Z.HomozygousAlternate, Z.PhasedHet01]),
      total_ed=2, gt=SummaryGtMetrics(2, 1, 1, 1, 0, 1), hap=SummaryMetrics(2, 1, 2, 1),
      basepair=SummaryMetrics(4, 2, 4, 2),
      truth=[VM(T, 1, 1), VM(T, 1, 0), VM(T, 1, 1)], query=[VM(Q, 1, 2), VM(Q, 1, 1)],
      seqs=("ACCGTTACCA", "ACGGTTCCCA", "ACCGCTACCA", "ACGGTTCCCA", "ACGGTTACCA"))

# :1158-1202 test_solve_compare_region_ambiguous (which hap gets A vs C is pinned)
_case("ambiguous",
      CompareRegion(0, _coords(0, 10),
                    [Variant.new_snv(0, 4, b"T", b"C")], [Z.HomozygousAlternate],
                    [Variant.new_snv(0, 4, b"T", b"C"), Variant.new_snv(0, 4, b"T", b"A")], [Z.PhasedHet01, Z.PhasedHet10]),
      total_ed=1, gt=SummaryGtMetrics(0, 1, 1, 1, 1, 0), hap=SummaryMetrics(1, 1, 1, 1),
      basepair=SummaryMetrics(3, 1, 3, 1),
      truth=[VM(T, 2, 1)], query=[VM(Q, 1, 1), VM(Q, 0, 1)],
      seqs=("ACCGTTACCA", "ACCGCTACCA", "ACCGCTACCA", "ACCGATACCA", "ACCGCTACCA"))

# :1204-1247 test_solve_compare_region_skip_variants (incl. per-type basepair)
_sv = [Variant.new_deletion(0, 3, b"GT", b"G"), Variant.new_snv(0, 4, b"T", b"G")]
_sz = [Z.PhasedHet01, Z.HomozygousAlternate]
_case("skip_variants",
      CompareRegion(0, _coords(0, 10), list(_sv), list(_sz), list(_sv), list(_sz)),
      total_ed=0, gt=SummaryGtMetrics(1, 1, 1, 1, 1, 1), hap=SummaryMetrics(2, 1, 2, 1),
      basepair=SummaryMetrics(4, 2, 4, 2),
      type_basepair={VariantType.Snv: SummaryMetrics(3, 1, 3, 1), VariantType.Deletion: SummaryMetrics(2, 0, 2, 0)},
      truth=[VM(T, 1, 1), VM(T, 2, 1)],
      query=[VariantMetrics.toggle_source(VM(T, 1, 1)), VariantMetrics.toggle_source(VM(T, 2, 1))],
      seqs=("ACCGTTACCA", "ACCGGTACCA", "ACCGTACCA", "ACCGGTACCA", "ACCGTACCA"))


# --- merge: src/merge_solver.rs:242-348 + doctest :9-47 ----------------------
def _snv5():
    return [Variant.new_snv(0, 5, b"T", b"C")]


HOM, H01 = Z.HomozygousAlternate, Z.PhasedHet01
MERGE_CASES = [
    # test_exact_mode :243-272
    ("exact_identical", MultiRegion(0, _coords(0, 25), [_snv5(), _snv5(), _snv5()], [[HOM], [HOM], [HOM]]),
     MergeConfig(), MergeClassification.BasepairIdentical),
    ("exact_different", MultiRegion(0, _coords(0, 25), [_snv5(), _snv5(), _snv5()], [[HOM], [HOM], [H01]]),
     MergeConfig(), MergeClassification.Different),
    # test_noconflict_mode :275-310
    ("noconflict", MultiRegion(0, _coords(0, 25), [_snv5(), _snv5(), []], [[HOM], [HOM], []]),
     MergeConfig(no_conflict_enabled=True), MergeClassification.NoConflict([0, 1])),
    ("noconflict_different", MultiRegion(0, _coords(0, 25), [_snv5(), _snv5(), _snv5()], [[HOM], [HOM], [H01]]),
     MergeConfig(no_conflict_enabled=True), MergeClassification.Different),
    # test_majority_mode :313-348
    ("majority", MultiRegion(0, _coords(0, 25), [_snv5(), _snv5(), []], [[HOM], [HOM], []]),
     MergeConfig(majority_voting_enabled=True), MergeClassification.MajorityAgree([0, 1])),
    ("majority_different", MultiRegion(0, _coords(0, 25), [_snv5(), [], _snv5()], [[HOM], [], [H01]]),
     MergeConfig(majority_voting_enabled=True), MergeClassification.Different),
    # doctest :9-47
    ("doctest_different", MultiRegion(0, _coords(0, 25), [_snv5(), _snv5(), _snv5()], [[HOM], [H01], [HOM]]),
     MergeConfig(), MergeClassification.Different),
    ("doctest_majority", MultiRegion(0, _coords(0, 25), [_snv5(), _snv5(), _snv5()], [[HOM], [H01], [HOM]]),
     MergeConfig(majority_voting_enabled=True), MergeClassification.MajorityAgree([0, 2])),
    # conflict selection (merge_solver.rs:192-194; no reference test -- checked against the oracle only)
    ("conflict_select", MultiRegion(0, _coords(0, 25), [_snv5(), _snv5(), _snv5()], [[HOM], [H01], [HOM]]),
     MergeConfig(conflict_selection=1), MergeClassification.ConflictSelection(1)),
]

# --- wfa_ed / edit_distance known answers --------------------------------------
# (a, b, ed): src/dwfa/dynamic_wfa.rs:289-403 and src/util/sequence_alignment.rs:57-116
_V1, _V2, _V3, _V4 = bytes([0, 1, 2, 4, 5]), bytes([0, 1, 3, 4, 5]), bytes([1, 2, 3, 5]), b""
_E1 = bytes([65] * 17 + [67, 65, 65, 65])
_E2 = bytes([65] * 10 + [67] + [65] * 6 + [67, 65, 65, 65])
_E3 = bytes([65] * 16 + [67, 65, 65, 65])
ED_CASES = [
    (b"", b"ACGT", 4), (b"ACGT", b"", 4),
    (b"ACGTACGTACGT", b"ACGTACGTACGT", 0),
    (b"ACGTACGTACGT", b"ACGTACCTACGT", 1),
    (b"ACGTACGTACGT", b"ACGTACIGTACGT", 1),
    (b"ACGTACGTACGT", b"ACGTACTACGT", 1),
    (b"ACGTACGTACGT", b"ACTACGCACGGGT", 4),
    (b"AACGGATCAAGCTTACCAGTATTTACGT", b"AACGGACAAAAGCTTACCTGTATTACGT", 5),
    (b"AA", b"ATA", 1),
    (b"ATTTTTTTTTTAAAAAAAAAA", b"AAAAAAAAAAA", 10),
    (b"ATTTTTTTTTTA", b"AA", 10),
    # src/util/sequence_alignment.rs:57-100 (test_edit_distance / test_wfa_ed)
    (_V1, _V1, 0), (_V1, _V2, 1), (_V1, _V3, 2), (_V1, _V4, 5),
    (_V2, _V2, 0), (_V2, _V3, 3), (_V2, _V4, 5),
    (_V3, _V3, 0), (_V3, _V4, 4), (_V4, _V4, 0),
    # src/util/sequence_alignment.rs:102-116 (test_edit_error_001)
    (_E1, _E3, 1), (_E2, _E3, 1), (_E3, _E1, 1), (_E3, _E2, 1),
]
