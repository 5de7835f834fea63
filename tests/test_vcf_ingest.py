"""VCF ingest: the oracle's restatement of parse_variant / parse_genotype / get_variant_type
(src/parsing/region_generation.rs:565-758) against expectations derived by hand from those rules (the reference has no tests
for its parsing functions, region_generation.rs:815-822: parity UNPINNED by reference vectors).  The device parser
(avk_vcf_parse) is compared with the oracle on the same text and on a synthetic call set written out as VCF in
tests/test_gpu_parity.py."""
import oracle_py as orc
from aardvark_b200 import abi

HEADER = b"##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\n"
LINES = [
    b"chr1\t101\t.\tA\tG\t50\tPASS\t.\tGT:DP\t0/1:30\t1/1:20",                 # unphased het SNV
    b"chr1\t105\trs1\tAT\tA\t.\t.\tDP=3\tGT\t1|0\t0|1",                        # phased het deletion: 1|0 -> PhasedHet10
    b"chr1\t110\t.\tC\tCAA,CA\t.\t.\t.\tDP:GT\t9:2|1\t1:1/2",                  # multi-ALT split: (2, Het10) then (1, Het01)
    b"chr1\t120\t.\tGTT\tGT,*\t.\t.\t.\tGT\t1/2\t./.",                         # trimmed GTT>GT -> GT>G; '*' dropped; S2 './.' = hom-ref
    b"chr1\t130\t.\tT\t<DEL>\t.\t.\tSVTYPE=DEL\tGT\t1/1\t0/0",                 # symbolic ALT dropped
    b"chr1\t140\t.\tTAAAA\tT\t.\t.\tSVTYPE=DEL;END=144\tGT\t1\t.",             # haploid = homozygous; SV deletion by tag
    b"chr2\t7\t.\tG\tGACAC\t.\t.\tTRID=chr2_7\tGT\t0|1\t1|1",                  # tandem-repeat expansion by tag
    b"chr2\t9\t.\tGAC\tG\t.\t.\tX;TRID=t9\tGT\t.|1\t0/1",                      # '.' allele = reference; TR contraction
    b"chr2\t20\t.\tACGT\tATTT\t.\t.\t.\tGT\t1/1\t1/1",                         # indel (no common suffix beyond T: ACG>ATT)
    b"chr2\t30\t.\tA\tT\t.\t.\tSVTYPE=BND\tGT\t0/1\t0/1",                      # BND dropped
    b"chr2\t40\t.\tA\tC\t.\t.\t.\tGT\t.\t0/0",                                 # GT '.' -> no-op
]
TEXT = HEADER + b"\n".join(LINES) + b"\n"


def test_oracle_vcf_parse_hand_derived():
    tab, err = orc.vcf_parse(TEXT, ["chr1", "chr2"], sample_index=0, enable_trimming=True)
    assert err == 0
    Z, T = abi, abi
    assert tab.records() == [
        (0, 100, b"A", b"G", Z.ZYG_UNPHASED_HET, T.VT_SNV, 1),
        (0, 104, b"AT", b"A", Z.ZYG_PHASED_HET10, T.VT_DELETION, 2),
        (0, 109, b"C", b"CA", Z.ZYG_PHASED_HET10, T.VT_INSERTION, 2),           # allele index 2 first (haplotype 1), raw max(1, 2)
        (0, 109, b"C", b"CAA", Z.ZYG_PHASED_HET01, T.VT_INSERTION, 3),
        (0, 119, b"GT", b"G", Z.ZYG_UNPHASED_HET, T.VT_DELETION, 3),            # GTT>GT trimmed to GT>G, raw_allele_space before trimming
        (0, 139, b"TAAAA", b"T", Z.ZYG_HOM_ALT, T.VT_SV_DELETION, 5),
        (1, 6, b"G", b"GACAC", Z.ZYG_PHASED_HET01, T.VT_TR_EXPANSION, 5),
        (1, 8, b"GAC", b"G", Z.ZYG_PHASED_HET01, T.VT_TR_CONTRACTION, 3),
        (1, 19, b"ACG", b"ATT", Z.ZYG_HOM_ALT, T.VT_INDEL, 4),
    ]
    # second sample, trimming off
    tab2, err = orc.vcf_parse(TEXT, ["chr1", "chr2"], sample_index=1, enable_trimming=False)
    assert err == 0
    r2 = tab2.records()
    assert r2[0] == (0, 100, b"A", b"G", Z.ZYG_HOM_ALT, T.VT_SNV, 1)
    assert r2[2:4] == [(0, 109, b"C", b"CAA", Z.ZYG_UNPHASED_HET, T.VT_INSERTION, 3), (0, 109, b"C", b"CA", Z.ZYG_UNPHASED_HET, T.VT_INSERTION, 2)]
    assert (1, 19, b"ACGT", b"ATTT", Z.ZYG_HOM_ALT, T.VT_INDEL, 4) in r2


def test_oracle_vcf_parse_errors():
    bad = {
        b"chr1\t5\t.\tA\tG\t.\t.\t.\tDP\t3\t4": 3,                              # no GT key: "Missing GT"
        b"chr1\t5\t.\tA\tG\t.\t.\t.\tGT\t0/1/1\t0/1": 4,                        # ploidy 3
        b"chr1\t5\t.\tA\tG\t.\t.\t.\tGT\t0/2\t0/1": 5,                          # ALT index out of range
        b"chr1\t5\t.\tA\tG\t.\t.\tSVTYPE=INV\tGT\t0/1\t0/1": 6,                 # unsupported SVTYPE
        b"chrX\t5\t.\tA\tG\t.\t.\t.\tGT\t0/1\t0/1": 7,                          # unknown contig
        b"chr1\t5\t.\tA\tGTT\t.\t.\tSVTYPE=DEL\tGT\t0/1\t0/1": 9,               # SvDeletion constructor: ALT longer than REF
    }
    for line, code in bad.items():
        tab, err = orc.vcf_parse(HEADER + LINES[0] + b"\n" + line + b"\n", ["chr1", "chr2"])
        assert tab is None and err == (3, code), (line, err)
