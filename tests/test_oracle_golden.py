"""Pins the CPU oracle (oracle/aardvark_oracle.cpp) against every golden vector the
reference's own unit tests / doctests hold for the hot path.  CPU only.
Each test names the reference test it restates (paths relative to the reference repo)."""
import json
import os

import pytest

import oracle_py as orc
from aardvark_b200 import abi
from aardvark_b200.batch import CompareOutputs, RegionBatch, seq_offsets
from aardvark_b200.results import unpack_compare, unpack_merge
from aardvark_b200.types import (Allele, CompareRegion, Coordinates, PhasedZygosity as Z, SummaryMetrics, Variant)
from golden_cases import COMPARE_CASES, ED_CASES, MERGE_CASES, MOCK_CHR1

HERE = os.path.dirname(os.path.abspath(__file__))
REF12 = b"ACGTACGTACGT"


# ---------------------------------------------------------------- DWFALite
def test_dwfa_new():  # src/dwfa/dynamic_wfa.rs:282-287
    d = orc.DWFALite()
    assert d.edit_distance() == 0
    assert d.wavefront() == [0]


def test_dwfa_empty():  # :289-304
    for a, b in ((b"", b"ACGT"), (b"ACGT", b"")):
        d = orc.DWFALite()
        assert d.finalize(a, b) == 0
        assert d.edit_distance() == 4
        assert d.wavefront() == [0, 1, 2, 3, 4, 4, 4, 4, 4]


@pytest.mark.parametrize("alt,ed", [  # :306-348 (char-by-char updates)
    (b"ACGTACGTACGT", 0), (b"ACGTACCTACGT", 1), (b"ACGTACIGTACGT", 1), (b"ACGTACTACGT", 1)])
def test_dwfa_incremental(alt, ed):
    d = orc.DWFALite()
    for l in range(len(alt)):
        assert d.update(REF12, alt[:l + 1]) == 0
        if ed == 0:
            assert d.edit_distance() == 0
    assert d.edit_distance() == ed


def test_dwfa_complex_001():  # :350-360
    alt = b"ACTACGCACGGGT"
    d = orc.DWFALite()
    for l in range(len(alt)):
        d.update(REF12, alt[:l + 1])
    d.finalize(REF12, alt)
    assert d.edit_distance() == 4


def test_dwfa_complex_002():  # :362-377 one-shot update
    d = orc.DWFALite()
    d.update(b"AACGGATCAAGCTTACCAGTATTTACGT", b"AACGGACAAAAGCTTACCTGTATTACGT")
    assert d.edit_distance() == 5


def test_dwfa_big_insertion():  # :379-390
    d = orc.DWFALite()
    d.finalize(b"AA", b"ATA")
    assert d.edit_distance() == 1


def test_dwfa_big_deletion():  # :392-403
    seq, alt = b"ATTTTTTTTTTAAAAAAAAAA", b"AAAAAAAAAAA"
    d = orc.DWFALite()
    for l in range(len(alt)):
        d.update(seq, alt[:l + 1])
    assert d.edit_distance() == len(seq) - len(alt)


def test_dwfa_required_finalize():  # :405-422
    seq, alt = b"ATTTTTTTTTTA", b"AA"
    d = orc.DWFALite()
    for l in range(len(alt)):
        d.update(seq, alt[:l + 1])
    assert d.edit_distance() == 1
    d.finalize(seq, alt)
    assert d.edit_distance() == len(seq) - len(alt)
    assert d.update(seq, alt) == 2  # AlreadyFinalized (:69-71)


def test_dwfa_cloning():  # :424-450
    seq, alt = b"AAAAAAA", b"AAACAAA"
    d = orc.DWFALite()
    d2 = d.clone()
    for l in range(len(alt)):
        d.update(seq, seq[:l + 1])
        d2.update(seq, alt[:l + 1])
        if seq[l] == alt[l]:
            assert d == d2
        else:
            assert not (d == d2)
            d2 = d.clone()
    assert d.edit_distance() == 0 and d2.edit_distance() == 0


def test_dwfa_big_early_termination():  # :452-468 (5929 vs 651 bp: <= 2 during updates, 5278 after finalize)
    g = json.load(open(os.path.join(HERE, "golden", "dwfa_big_early_termination.json")))
    c1, s23 = g["c1"].encode(), g["seq_23"].encode()
    d = orc.DWFALite()
    for i in range(len(c1)):
        d.update(s23, c1[:i + 1])
        assert d.edit_distance() <= 2
    assert d.edit_distance() == 2
    d.finalize(s23, c1)
    assert d.edit_distance() == 5278
    assert orc.wfa_ed(s23, c1) == 5278
    # path independence: one-shot update == incremental updates
    e = orc.DWFALite()
    e.update(s23, c1)
    assert e.edit_distance() == 2


def test_dwfa_max_edit_distance_leaves_incremented_ed():  # :146-149
    d = orc.DWFALite(0)
    assert d.update(b"ACGT", b"ACCT") == 1
    assert d.edit_distance() == 1


@pytest.mark.parametrize("a,b,ed", ED_CASES)
def test_wfa_ed_and_edit_distance(a, b, ed):  # src/util/sequence_alignment.rs:57-116 + dynamic_wfa.rs finalize cases
    assert orc.wfa_ed(a, b) == ed
    assert orc.edit_distance(a, b) == ed


# ---------------------------------------------------------------- HaplotypeDWFA
def test_haplotype_node():  # src/dwfa/haplotype_dwfa.rs:251-295
    h = orc.HaplotypeDWFA(3)
    v = Variant.new_snv(0, 6, b"G", b"A")
    h.extend_variant(REF12, True, v, Allele.Alternate)
    h.extend_variant(REF12, False, v, Allele.Alternate)
    v2 = Variant.new_snv(0, 7, b"T", b"A")
    h.extend_variant(REF12, False, v2, Allele.Alternate)
    assert h.edit_distance() == 0
    v3 = Variant.new_insertion(0, 8, b"A", b"TTTTTTTT")
    h.extend_variant(REF12, False, v3, Allele.Reference)
    assert h.edit_distance() == 0
    assert h.finalize_dwfa(REF12, 11) == 0
    assert h.edit_distance() == 1
    assert h.alleles(True) == [Allele.Alternate]
    assert h.alleles(False) == [Allele.Alternate, Allele.Alternate, Allele.Reference]
    assert h.sequence(True) == b"TACATACG"
    assert h.sequence(False) == b"TACAAACG"


def test_haplotype_node_incompatible():  # :297-332
    h = orc.HaplotypeDWFA(0)
    v = Variant.new_snv(0, 6, b"G", b"A")
    h.extend_variant(REF12, True, v, Allele.Alternate)
    h.extend_variant(REF12, False, v, Allele.Alternate)
    v2 = Variant.new_snv(0, 6, b"G", b"T")
    assert h.extend_variant(REF12, False, v2, Allele.Alternate) == 0  # not incorporated
    assert h.edit_distance() == 0
    assert h.total_variant_skip_distance() == 1
    h.finalize_dwfa(REF12, len(REF12))
    assert h.edit_distance() == 0
    assert h.total_variant_skip_distance() == 1
    assert h.total_cost() == 1


# ---------------------------------------------------------------- optimize_sequences
def _batch(ref, start, end, tv, tz, qv, qz):
    region = CompareRegion(0, Coordinates("c", start, end), tv, tz, qv, qz)
    return RegionBatch.from_compare_regions([region], {"c": 0})


INS = Variant.new_insertion(0, 4, b"A", b"AC")
SNV5G = Variant.new_snv(0, 5, b"C", b"G")
SNV8 = Variant.new_snv(0, 8, b"A", b"G")


def test_optimize_query_sequences_001():  # src/query_optimizer.rs:532-571
    tz = [Z.PhasedHet10, Z.PhasedHet01]
    qz = [Z.PhasedHet10, Z.PhasedHet01, Z.HomozygousAlternate]
    st, res = orc.optimize_sequences(_batch(REF12, 0, 12, [INS, SNV5G], tz, [INS, SNV5G, SNV8], qz), REF12)
    assert st == 0
    r = res[0]
    assert (r["ed1"], r["ed2"]) == (1, 1)
    assert r["truth_seq1"] == b"ACGTACCGTACGT" and r["truth_seq2"] == b"ACGTAGGTACGT"
    assert r["truth_zygosity"] == [int(z) for z in tz]
    assert r["query_seq1"] == b"ACGTACCGTGCGT" and r["query_seq2"] == b"ACGTAGGTGCGT"
    assert r["query_zygosity"] == [int(z) for z in qz]


def test_optimize_query_sequences_all_fn():  # :573-599
    st, res = orc.optimize_sequences(_batch(REF12, 0, 12, [INS, SNV5G], [Z.PhasedHet10, Z.PhasedHet01], [], []), REF12)
    r = res[0]
    assert (r["ed1"], r["ed2"]) == (1, 1)
    assert r["query_seq1"] == REF12 and r["query_seq2"] == REF12
    assert r["query_zygosity"] == []


def test_optimize_query_sequences_multiallelic():  # :601-632 and doctest :10-51
    shared = [Variant.new_snv(0, 5, b"C", b"A"), Variant.new_snv(0, 5, b"C", b"G")]
    tz = [Z.PhasedHet10, Z.PhasedHet01]
    qz = [Z.UnphasedHeterozygous, Z.UnphasedHeterozygous]
    st, res = orc.optimize_sequences(_batch(REF12, 0, 12, shared, tz, shared, qz), REF12)
    r = res[0]
    assert (r["ed1"], r["ed2"]) == (0, 0)
    assert r["query_seq1"] == b"ACGTAAGTACGT" and r["query_seq2"] == b"ACGTAGGTACGT"
    assert r["query_zygosity"] == [int(z) for z in tz]


def test_optimize_query_sequences_incompatible():  # :634-665
    shared = [Variant.new_snv(0, 5, b"C", b"A"), Variant.new_snv(0, 5, b"C", b"G")]
    hz = [Z.HomozygousAlternate, Z.HomozygousAlternate]
    st, res = orc.optimize_sequences(_batch(REF12, 0, 12, shared, hz, shared, hz), REF12)
    r = res[0]
    assert (r["ed1"], r["ed2"]) == (0, 0)
    assert (r["truth_vs1"], r["truth_vs2"], r["query_vs1"], r["query_vs2"]) == (1, 1, 1, 1)
    assert r["query_zygosity"] == [int(z) for z in hz]


def test_comparison_node():  # src/query_optimizer.rs:503-530 (a ComparisonNode is two HaplotypeDWFAs)
    h1, h2 = orc.HaplotypeDWFA(0), orc.HaplotypeDWFA(0)
    for v, a1, a2 in ((INS, Allele.Alternate, Allele.Reference), (SNV5G, Allele.Reference, Allele.Alternate),
                      (SNV8, Allele.Alternate, Allele.Alternate)):
        h1.extend_variant(REF12, False, v, a1)
        h2.extend_variant(REF12, False, v, a2)
    h1.finalize_dwfa(REF12, len(REF12))
    h2.finalize_dwfa(REF12, len(REF12))
    assert h1.total_cost() + h2.total_cost() == 4   # 1 INS, 1 SNP, 1 hom SNP = 1+1+2
    assert h1.sequence(False) == b"ACGTACCGTGCGT"
    assert h1.alleles(False) == [Allele.Alternate, Allele.Reference, Allele.Alternate]
    assert h2.sequence(False) == b"ACGTAGGTGCGT"
    assert h2.alleles(False) == [Allele.Reference, Allele.Alternate, Allele.Alternate]


# ---------------------------------------------------------------- optimize_gt_alleles
def _abatch(tv, ta, qv, qa):
    return _batch(REF12, 0, 12, tv, ta, qv, qa)  # zygosity column carries Allele codes


A = Allele


def test_optimize_gt_alleles_match():  # src/exact_gt_optimizer.rs:525-548 and doctest :8-41
    st, t, q, ne = orc.optimize_gt_alleles(_abatch([SNV5G], [A.Alternate], [SNV5G], [A.Alternate]), REF12)
    assert (st, ne, t, q) == (0, 0, [A.Alternate], [A.Alternate])


def test_optimize_gt_alleles_all_fn():  # :550-573
    st, t, q, ne = orc.optimize_gt_alleles(_abatch([SNV5G], [A.Alternate], [], []), REF12)
    assert (st, ne, t, q) == (0, 1, [A.Reference], [])


def test_optimize_gt_alleles_all_fp():  # :575-598
    st, t, q, ne = orc.optimize_gt_alleles(_abatch([], [], [SNV5G], [A.Alternate]), REF12)
    assert (st, ne, t, q) == (0, 1, [], [A.Reference])


def test_optimize_gt_alleles_diff_rep():  # :600-629
    tv = [Variant.new_indel(0, 5, b"CGT", b"GGG")]
    qv = [Variant.new_snv(0, 5, b"C", b"G"), Variant.new_snv(0, 7, b"T", b"G")]
    st, t, q, ne = orc.optimize_gt_alleles(_abatch(tv, [A.Alternate], qv, [A.Alternate, A.Alternate]), REF12)
    assert (st, ne, t, q) == (0, 0, [A.Alternate], [A.Alternate, A.Alternate])


def test_exact_match_node():  # :494-523: ED short-circuits to 1 under max ED 0
    h = orc.HaplotypeDWFA(0, 0)
    h.extend_variant(REF12, False, INS, A.Alternate)
    h.extend_variant(REF12, False, SNV5G, A.Reference)
    h.extend_variant(REF12, False, SNV8, A.Alternate)
    assert h.finalize_dwfa(REF12, len(REF12)) == 1   # MaxEditDistance, tolerated (:425-433)
    assert h.edit_distance() == 1
    assert h.sequence(True) == REF12
    assert h.alleles(True) == []
    assert h.sequence(False) == b"ACGTACCGTGCGT"
    assert h.alleles(False) == [A.Alternate, A.Reference, A.Alternate]


# ---------------------------------------------------------------- waffle solver pieces
def _hb(start, end, vs, zs):
    return RegionBatch.from_compare_regions([CompareRegion(0, Coordinates("c", start, end), vs, zs, [], [])], {"c": 0})


def test_generate_haplotype_sequence_001():  # src/waffle_solver.rs:814-838
    b = _hb(5, 15, [Variant.new_snv(0, 10, b"G", b"C")], [Z.PhasedHet10])
    assert orc.generate_haplotype_sequence(b, MOCK_CHR1, 0) == (0, b"TACCACGACT", 0)
    assert orc.generate_haplotype_sequence(b, MOCK_CHR1, 1) == (0, b"TACCAGGACT", 0)


def test_generate_haplotype_sequence_002():  # :840-866
    vs = [Variant.new_deletion(0, 3, b"GT", b"G"), Variant.new_snv(0, 4, b"T", b"G")]
    b = _hb(0, 10, vs, [Z.PhasedHet10, Z.PhasedHet01])
    assert orc.generate_haplotype_sequence(b, MOCK_CHR1, 0) == (0, b"ACCGTACCA", 0)
    assert orc.generate_haplotype_sequence(b, MOCK_CHR1, 1) == (0, b"ACCGGTACCA", 0)


def test_generate_haplotype_sequence_conflict():  # :868-894
    vs = [Variant.new_deletion(0, 3, b"GT", b"G"), Variant.new_snv(0, 4, b"T", b"G")]
    b = _hb(0, 10, vs, [Z.PhasedHet01, Z.HomozygousAlternate])
    assert orc.generate_haplotype_sequence(b, MOCK_CHR1, 0) == (0, b"ACCGGTACCA", 0)
    assert orc.generate_haplotype_sequence(b, MOCK_CHR1, 1) == (0, b"ACCGTACCA", 1)


def test_perform_basepair_compare():  # :1249-1298
    f = orc.perform_basepair_compare
    assert f(b"ACGTACGT", b"ACGTACGT", b"ACCTAGGT") == (0, 0, 0, 4)
    assert f(b"ACGTACGT", b"ACCTAGGT", b"ACGTACGT") == (0, 4, 0, 0)
    assert f(b"ACGTACGT", b"ACCTAGGT", b"ACCTAGGT") == (4, 0, 4, 0)
    assert f(b"TAT", b"TCT", b"TACT") == (1, 1, 1, 1)
    assert f(b"TAT", b"TAAAAT", b"TAAAT") == (4, 2, 4, 0)
    assert f(b"TAT", b"TAAAAT", b"TAAAAAT") == (6, 0, 6, 2)


def check_compare_case(bench, exp):
    """Shared with the GPU parity test: the assertions of waffle_solver.rs:896-1247."""
    assert bench.total_ed() == exp["total_ed"]
    jm = bench.group_metrics().joint_metrics()
    assert jm.gt() == exp["gt"]
    assert jm.hap() == exp["hap"]
    assert jm.basepair() == exp["basepair"]
    for vt, sm in exp.get("type_basepair", {}).items():
        assert bench.group_metrics().variant_metrics()[vt].basepair() == sm
    assert bench.truth_variant_data() == exp["truth"]
    assert bench.query_variant_data() == exp["query"]
    sb = bench.sequence_bundle()
    assert (sb.ref_seq, sb.truth_seq1, sb.truth_seq2, sb.query_seq1, sb.query_seq2) == exp["seqs"]


@pytest.mark.parametrize("name,region,exp", COMPARE_CASES, ids=[c[0] for c in COMPARE_CASES])
def test_solve_compare_region(name, region, exp):  # src/waffle_solver.rs:896-1247, doctests lib.rs:9-63
    batch = RegionBatch.from_compare_regions([region], {"mock_chr1": 0})
    off, plen = seq_offsets(batch)
    out = orc.compare_batch(batch, [MOCK_CHR1], abi.CompareCfg(50, 0, 1, 0), seq_off=off, seq_pool_len=plen)
    bench = unpack_compare(batch, out)[0]
    check_compare_case(bench, exp)
    # the reduction equals the single region (writers/summary.rs:146-158)
    assert (out.totals == out.region_metrics[0]).all()
    assert int(out.solved_blocks[0]) == 1 and int(out.error_blocks[0]) == 0


def test_compare_all_cases_one_batch():
    """All 8 reference cases as one batch == BASELINE config 1 (the CPU-runnable reference case)."""
    regions = []
    for i, (_, r, _) in enumerate(COMPARE_CASES):
        regions.append(CompareRegion(i, r.coordinates, r.truth_variants, r.truth_zygosity, r.query_variants, r.query_zygosity))
    batch = RegionBatch.from_compare_regions(regions, {"mock_chr1": 0})
    off, plen = seq_offsets(batch)
    out = orc.compare_batch(batch, [MOCK_CHR1], abi.CompareCfg(50, 0, 1, 0), seq_off=off, seq_pool_len=plen, n_threads=2)
    for bench, (_, _, exp) in zip(unpack_compare(batch, out), COMPARE_CASES):
        check_compare_case(bench, exp)
    assert (out.totals == out.region_metrics.sum(axis=0)).all()
    assert int(out.solved_blocks[0]) == len(regions)


def test_exact_shortcut():  # src/waffle_solver.rs:171-199, 534-601
    name, region, exp = COMPARE_CASES[1]  # double_single: exact match through different representations
    batch = RegionBatch.from_compare_regions([region], {"mock_chr1": 0})
    out = orc.compare_batch(batch, [MOCK_CHR1], abi.CompareCfg(50, 1, 0, 0))
    b = unpack_compare(batch, out)[0]
    jm = b.group_metrics().joint_metrics()
    assert jm.gt() == exp["gt"] and jm.hap() == exp["hap"] and jm.basepair() == exp["basepair"]
    assert jm.record_bp() == SummaryMetrics(0, 0, 0, 0)   # shortcut returns before add_record_basepair_stats
    assert b.truth_variant_data() == exp["truth"] and b.query_variant_data() == exp["query"]


# ---------------------------------------------------------------- merge solver
@pytest.mark.parametrize("name,region,cfg,expected", MERGE_CASES, ids=[c[0] for c in MERGE_CASES])
def test_solve_merge_region(name, region, cfg, expected):  # src/merge_solver.rs:242-348, doctest :9-47
    batch = RegionBatch.from_multi_regions([region], {"mock_chr1": 0})
    ccfg = abi.MergeCfg(cfg.max_branch_factor, int(cfg.no_conflict_enabled), int(cfg.majority_voting_enabled),
                        -1 if cfg.conflict_selection is None else cfg.conflict_selection)
    out = orc.merge_batch(batch, [MOCK_CHR1], ccfg)
    assert unpack_merge(batch, out)[0].merge_classification == expected


def test_variant_length_delta():  # src/merge_solver.rs:350-370
    vs = [Variant.new_snv(0, 10, b"A", b"C"), Variant.new_deletion(0, 12, b"ACGTACGT", b"A"),
          Variant.new_insertion(0, 25, b"A", b"ACC")]
    zs = [Z.HomozygousAlternate, Z.PhasedHet01, Z.HomozygousAlternate]
    for i, d in enumerate((0, -7, 4)):
        assert orc.variant_delta_length(_hb(0, 40, vs[i:i + 1], zs[i:i + 1])) == d
    assert orc.variant_delta_length(_hb(0, 40, vs, zs)) == -3


def test_region_builder_oracle_matches_host_builder():
    """the reference has no tests for its region builder (parity unpinned); the C++ restatement is
    pinned against the generator's host builder on compare and merge call sets and on hand-made corner cases."""
    import numpy as np
    from aardvark_b200 import abi, synth
    from aardvark_b200.batch import CallSets

    def same(a, b):
        for f in ("region_id", "contig", "start", "end", "var_off", "position", "variant_type", "zygosity", "raw_allele_space",
                  "allele_off", "a0_len", "a1_len"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), f
        n = int(a.a0_len.sum() + a.a1_len.sum())
        assert np.array_equal(a.allele_pool[:n], b.allele_pool[:n])

    ref, inputs = synth.callsets_compare(300_000, synth.SynthParams(n_variants=700), seed=81)
    for flank in (0, 50, 1000):
        same(orc.build_regions(CallSets(inputs), len(ref), flank), synth.cluster_regions(inputs, len(ref), flank))
    ref5, sets, flank5 = synth.callsets_merge(150_000, 400, n_sets=5, seed=82)
    same(orc.build_regions(CallSets(sets), len(ref5), flank5, first_region_id=7),
         synth.cluster_regions(sets, len(ref5), flank5, first_region_id=7))
    small = bytes(synth.ACGT[np.random.default_rng(5).integers(0, 4, size=600)])
    rec = lambda p, a0, a1: synth.make_rec(p, a0, a1, abi.ZYG_HOM_ALT)
    A = [rec(10, small[10:11], b"TT"), rec(10, small[10:11], b"AAAC"), rec(300, small[300:303], small[300:301]),
         rec(598, small[598:600], b"A"), rec(599, small[599:600] + b"A", b"C")]
    B = [rec(10, small[10:11], b"CC"), rec(61, small[61:62], b"AG"), rec(352, small[352:353], b"TG")]
    for sets3, flank in (([A, B], 50), ([A, [], B], 50), ([[], []], 50), ([B, A], 0), ([A, B], 5000)):
        same(orc.build_regions(CallSets(sets3), len(small), flank), synth.cluster_regions(sets3, len(small), flank))
