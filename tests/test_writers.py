"""Writers.  Golden: the example block of the reference's docs/compare.md:79-87 (tests/golden/summary_doc_example.tsv,
copied verbatim; the last line's two empty trailing fields restored) -- real rows of `aardvark compare`, which pin the row
layout, the Joint* categories and the f64 recall / precision / F1 values and their shortest-round-trip formatting.  The
reference has no unit tests for its writers.  These functions are host-only: they run without a GPU."""
import os

import numpy as np

from aardvark_b200 import abi
from aardvark_b200.batch import CompareOutputs, RegionBatch
from aardvark_b200.types import Coordinates, CompareRegion, PhasedZygosity, Variant, VariantType
from aardvark_b200.writers import SummaryWriter, vcf_record_lines

HERE = os.path.dirname(os.path.abspath(__file__))


def test_summary_rows_match_the_reference_docs_example():
    gold = open(os.path.join(HERE, "golden", "summary_doc_example.tsv")).read().splitlines()
    rows = {(f[1], f[4]): [int(x) for x in (f[6], f[7], f[9], f[10])] + [int(x) if x else 0 for x in f[14:16]] for f in (l.split("\t") for l in gold[1:])}
    tot = np.zeros((abi.N_GROUPS, abi.N_METRICS), dtype=np.uint64)
    col = {"GT": abi.M_GT, "BASEPAIR": abi.M_BASEPAIR}
    for (metric, vt), vals in rows.items():
        g = {"ALL": 0, "Snv": 1 + abi.VT_SNV, "JointIndel": 1 + abi.VT_INDEL}[vt]      # the whole JointIndel row as one member type
        tot[g, col[metric]:col[metric] + 4] = vals[:4]
        if metric == "GT":
            tot[g, abi.M_GT + 4:abi.M_GT + 6] = vals[4:6]
    text = SummaryWriter("compare", ["GT", "BASEPAIR"]).summary_text(tot).splitlines()
    assert text[0] == gold[0]
    ours = {(l.split("\t")[1], l.split("\t")[4]): l for l in text[1:]}
    for g in gold[1:]:
        f = g.split("\t")
        assert ours[(f[1], f[4])] == g
    # the member type's own row is there too, in VariantType order between the ALL row and the joint rows
    assert [l.split("\t")[4] for l in text[1:5]] == ["ALL", "Snv", "Indel", "JointIndel"]


def test_summary_empty_metrics_strata_and_csv():
    tot = np.zeros((abi.N_GROUPS, abi.N_METRICS), dtype=np.uint64)
    tot[0, abi.M_HAP:abi.M_HAP + 4] = [0, 0, 3, 1]                                  # no truth entries: recall and F1 undefined
    tot[1 + abi.VT_SV_DELETION, abi.M_HAP:abi.M_HAP + 4] = [0, 0, 3, 1]
    tot[1 + abi.VT_TR_EXPANSION, abi.M_HAP:abi.M_HAP + 4] = [1, 99999, 0, 0]          # recall 1e-5: still plain decimals
    w = SummaryWriter("x", ["HAP"], strat_labels=["easy", "hard"])
    strat = np.stack([tot, np.zeros_like(tot)])
    lines = w.summary_text(tot, strat, csv=True).splitlines()
    assert lines[1] == "x,HAP,ALL,ALL,ALL,0,0,0,4,3,1,,0.75,,,"
    assert lines[2] == "x,HAP,ALL,ALL,SvDeletion,0,0,0,4,3,1,,0.75,,,"
    assert lines[3] == "x,HAP,ALL,ALL,TrExpansion,100000,1,99999,0,0,0,0.00001,,,,"
    assert lines[4].startswith("x,HAP,ALL,ALL,JointStructuralVariant,") and lines[5].startswith("x,HAP,ALL,ALL,JointTandemRepeat,")
    assert lines[6] == "x,HAP,easy,ALL,ALL,0,0,0,4,3,1,,0.75,,,"                      # region_label = the stratum, filter stays ALL (:207-214)
    assert lines[-1] == "x,HAP,hard,ALL,ALL,0,0,0,0,0,0,,,,,"                           # an empty stratum still gets its ALL row
    assert len(lines) == 1 + 5 + 5 + 1


def test_vcf_record_lines():
    v = lambda pos, a0, a1: Variant(0, VariantType.Snv if len(a0) == len(a1) == 1 else VariantType.Deletion, pos, a0, a1, max(len(a0), len(a1)))
    region = CompareRegion(41, Coordinates("chr7", 90, 130), [v(100, b"A", b"G"), v(110, b"CTT", b"C")],
                           [PhasedZygosity.PhasedHet01, PhasedZygosity.HomozygousAlternate], [v(100, b"A", b"G")], [PhasedZygosity.UnphasedHeterozygous])
    batch = RegionBatch.from_compare_regions([region], {"chr7": 0})
    out = CompareOutputs(batch)
    out.var_class[:3] = [abi.CLASS_TP, abi.CLASS_FN, abi.CLASS_TP]
    out.var_expected[:3] = [1, 2, 1]
    out.var_observed[:3] = [1, 0, 1]
    assert vcf_record_lines(batch, 0, ["chr7"], out) == ("chr7\t101\t.\tA\tG\t.\t.\t.\tGT:BD:EA:OA:RI\t0|1:TP:1:1:41\n"
                                                        "chr7\t111\t.\tCTT\tC\t.\t.\t.\tGT:BD:EA:OA:RI\t1/1:FN:2:0:41\n")
    assert vcf_record_lines(batch, 1, ["chr7"], out) == "chr7\t101\t.\tA\tG\t.\t.\t.\tGT:BD:EA:OA:RI\t0/1:TP:1:1:41\n"


# ---- aardvark merge writers -------------------------------------------------------------------------------------------
# Goldens: the example blocks of the reference's docs/merge.md:70-78 (passing.vcf.gz body lines) and :97-121 (merge summary
# rows), copied verbatim into tests/golden/merge_doc_example.vcf and merge_summary_doc_example.tsv.  They pin the record
# layout (INFO SOURCES / MR, FORMAT GT:RI), which input's labels are listed, and the row order of the summary (the derived Ord
# of MergeClassification: different < no_conflict_* < majority_* < identical, index lists lexicographic).

def _merge_batch(regions_spec, k, chrom="chr1"):
    """regions_spec: [(region_id, start, end, [[(pos, ref, alt, zyg)] per input])] -> (RegionBatch, contig names)"""
    from aardvark_b200.types import MultiRegion
    regs = []
    for rid, s, e, per_input in regions_spec:
        vs = [[Variant(i, _vt(r, a), p, r, a) for (p, r, a, _z) in lst] for i, lst in enumerate(per_input)]
        zs = [[z for (_p, _r, _a, z) in lst] for lst in per_input]
        regs.append(MultiRegion(rid, Coordinates(chrom, s, e), vs, zs))
    return RegionBatch.from_multi_regions(regs, {chrom: 0}), [chrom]


def _vt(r, a):
    if len(r) == 1 and len(a) == 1:
        return VariantType.Snv
    if len(r) == 1:
        return VariantType.Insertion
    if len(a) == 1:
        return VariantType.Deletion
    return VariantType.Indel


def _merge_out(batch, classes):
    """classes: [(AVK_MERGE_* code, [indices])] per region"""
    from aardvark_b200.batch import MergeOutputs
    out = MergeOutputs(batch)
    out.status[:] = 0
    for r, (c, idx) in enumerate(classes):
        out.classification[r] = c
        out.n_indices[r] = len(idx)
        out.indices[r, :len(idx)] = idx
    return out


DIFFERENT, NO_CONFLICT, MAJORITY, CONFLICT_SELECT, IDENTICAL = range(5)
HOM = PhasedZygosity.HomozygousAlternate


def test_merge_records_match_the_reference_docs_example():
    from aardvark_b200.writers import VariantMerger
    gold = open(os.path.join(HERE, "golden", "merge_doc_example.vcf")).read()
    lines = [l.split("\t") for l in gold.splitlines()]
    # regions 0..7 as the RI column groups them; the passing inputs as SOURCES lists them (pb = 0, ilmn = 1, ont = 2)
    label_ix = {"pb": 0, "ilmn": 1, "ont": 2}
    by_region = {}
    for f in lines:
        by_region.setdefault(int(f[9].split(":")[1]), []).append(f)
    spec, classes = [], []
    for rid in sorted(by_region):
        fs = by_region[rid]
        info = dict(kv.split("=") for kv in fs[0][7].split(";"))
        srcs = [label_ix[s] for s in info["SOURCES"].split(",")]
        code = {"identical": IDENTICAL, "no_conflict": NO_CONFLICT, "majority": MAJORITY}[info["MR"]]
        variants = [(int(f[1]) - 1, f[3].encode(), f[4].encode(), HOM) for f in fs]
        decoy = [(int(fs[0][1]) - 1, fs[0][3].encode(), b"N", HOM)]          # what a non-passing input holds: never written
        per_input = [variants if (code == IDENTICAL or i in srcs) else (decoy if code == MAJORITY else []) for i in range(3)]
        spec.append((rid, int(fs[0][1]) - 51, int(fs[-1][1]) + 50, per_input))
        classes.append((code, [] if code == IDENTICAL else srcs))
    # a failed region and an unsolved one in between: neither reaches the VCF
    spec.insert(3, (100, 90000, 90100, [[(90050, b"A", b"C", HOM)], [(90050, b"A", b"G", HOM)], []]))
    classes.insert(3, (DIFFERENT, []))
    spec.insert(5, (101, 106000, 106100, [[(106050, b"A", b"C", HOM)], [], []]))
    classes.insert(5, (NO_CONFLICT, [0]))
    batch, names = _merge_batch(spec, 3)
    out = _merge_out(batch, classes)
    out.status[5] = 7                                                        # a region whose solve returned an error
    vm = VariantMerger(batch, out, names, ["pb", "ilmn", "ont"])
    assert vm.passing_records() == gold
    bed = vm.regions_bed(True).splitlines()
    assert bed[0] == "chr1\t38181\t38282\tno_conflict_0" and bed[3] == "chr1\t105228\t105329\tno_conflict_3" and len(bed) == 8
    assert [l.split("\t")[3] for l in bed] == ["no_conflict_0", "identical_1", "no_conflict_2", "no_conflict_3", "identical_4", "majority_5", "identical_6", "identical_7"]
    assert vm.regions_bed(False) == "chr1\t90000\t90100\tdifferent_100\n"
    assert vm.passing_records(0, 2).splitlines() == gold.splitlines()[:2]


def test_merge_summary_rows_follow_the_reference_docs_example():
    from aardvark_b200.writers import VariantMerger
    gold = open(os.path.join(HERE, "golden", "merge_summary_doc_example.tsv")).read().splitlines()
    rows = [l.split("\t") for l in gold[1:]]
    # one region per merge reason of the example, with as many SNVs per input as needed to make the same cells non-zero
    reasons = []
    for f in rows:
        if f[0] not in reasons:
            reasons.append(f[0])
    spec, classes = [], []
    want = {}
    for n, reason in enumerate(reversed(reasons)):                           # fed in reverse: the writer sorts
        parts = reason.split("_")
        idx = [int(x) for x in parts if x.isdigit()]
        code = {"different": DIFFERENT, "no": NO_CONFLICT, "majority": MAJORITY, "identical": IDENTICAL}[parts[0]]
        present = [i for i in range(3) if any(f[0] == reason and int(f[2]) == i for f in rows)]
        per_input = [[(1000 * n + 10 * j + 5, b"A", b"C", HOM) for j in range(i + 1 + n % 2)] if i in present else [] for i in range(3)]
        spec.append((n, 1000 * n, 1000 * n + 100, per_input))
        classes.append((code, idx))
        for i in present:
            passing = code == IDENTICAL or i in idx
            want[(reason, i)] = (len(per_input[i]), 0) if passing else (0, len(per_input[i]))
    batch, names = _merge_batch(spec, 3)
    out = _merge_out(batch, classes)
    text = VariantMerger(batch, out, names, ["pb", "ilmn", "ont"]).summary_text().splitlines()
    assert text[0] == gold[0]
    ours = [l.split("\t") for l in text[1:]]
    assert [f[:4] for f in ours] == [f[:4] for f in rows]                      # same keys, same labels, same ORDER
    for f, g in zip(ours, rows):
        assert (int(f[4]), int(f[5])) == want[(f[0], int(f[2]))]
        assert (int(f[4]) > 0, int(f[5]) > 0) == (int(g[4]) > 0, int(g[5]) > 0)    # the example's zero pattern
    # two regions with the same key add up; types sort in VariantType order inside a reason; conflict_select; csv quoting
    spec2 = [(0, 0, 100, [[(10, b"A", b"C", HOM), (20, b"AT", b"A", HOM)], [(10, b"A", b"G", HOM)]]),
             (1, 200, 300, [[(210, b"A", b"C", HOM)], [(210, b"A", b"CT", HOM)]]),
             (2, 400, 500, [[(410, b"A", b"C", HOM)], [(410, b"A", b"T", HOM)]])]
    b2, names = _merge_batch(spec2, 2)
    o2 = _merge_out(b2, [(CONFLICT_SELECT, [1]), (CONFLICT_SELECT, [1]), (DIFFERENT, [])])
    t2 = VariantMerger(b2, o2, names, ["a,b", 'q"x']).summary_text(csv=True).splitlines()
    assert t2 == ["merge_reason,variant_type,vcf_index,vcf_label,pass_variants,fail_variants",
                  'different,Snv,0,"a,b",0,1', 'different,Snv,1,"q""x",0,1',
                  'conflict_select_1,Snv,0,"a,b",0,2', 'conflict_select_1,Snv,1,"q""x",1,0',
                  'conflict_select_1,Insertion,1,"q""x",1,0', 'conflict_select_1,Deletion,0,"a,b",0,1']
    assert VariantMerger(b2, o2, names, ["a", "b"]).passing_records().splitlines() == [
        "chr1\t11\t.\tA\tG\t.\t.\tSOURCES=b;MR=conflict_select\tGT:RI\t1/1:0", "chr1\t211\t.\tA\tCT\t.\t.\tSOURCES=b;MR=conflict_select\tGT:RI\t1/1:1"]
