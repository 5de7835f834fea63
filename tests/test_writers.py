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
