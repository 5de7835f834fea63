"""Host-side mirror of the reference's problem/result types for the hot path.

Names, argument meaning and error behaviour follow the reference so that the
parity tests read like the reference's own unit tests:

  Variant / VariantType         src/data_types/variants.rs:6-31, 73-355
  PhasedZygosity / Allele       src/data_types/phase_enums.rs
  Coordinates                   src/data_types/coordinates.rs
  CompareRegion / MultiRegion   src/data_types/compare_region.rs:13-60, multi_region.rs:9-50
  CompareConfig / MergeConfig   src/waffle_solver.rs:94-115, src/merge_solver.rs:62-84
  SummaryMetrics(+Gt), GroupMetrics, VariantMetrics, CompareBenchmark,
  MergeClassification, MergeBenchmark   src/data_types/*.rs

These are plain containers; all arithmetic happens in the CUDA library.
"""
from dataclasses import dataclass, field
from enum import IntEnum
from typing import List, Optional

from . import abi


class VariantType(IntEnum):
    Snv = 0
    Insertion = 1
    Deletion = 2
    Indel = 3
    SvInsertion = 4
    SvDeletion = 5
    SvDuplication = 6
    SvInversion = 7
    SvBreakend = 8
    TrContraction = 9
    TrExpansion = 10
    Unknown = 11


class PhasedZygosity(IntEnum):
    Unknown = 0
    HomozygousReference = 1
    UnphasedHeterozygous = 2
    PhasedHet01 = 3
    PhasedHet10 = 4
    HomozygousAlternate = 5

    def to_allele_count(self) -> int:
        return {0: 0, 1: 0, 2: 1, 3: 1, 4: 1, 5: 2}[int(self)]


class Allele(IntEnum):
    Unknown = 0
    Reference = 1
    Alternate = 2


class Classification(IntEnum):
    Unknown = 0
    TruePositive = 1
    FalseNegative = 2
    FalsePositive = 3


class VariantError(ValueError):
    pass


@dataclass
class Variant:
    """src/data_types/variants.rs:73-91 (constructors :93-355 enforce the same length rules)."""
    vcf_index: int
    variant_type: VariantType
    position: int
    allele0: bytes
    allele1: bytes
    raw_allele_space: int = 0

    def __post_init__(self):
        if self.raw_allele_space == 0:
            self.raw_allele_space = max(len(self.allele0), len(self.allele1))

    @staticmethod
    def _chk(a0, a1):
        if len(a0) == 0:
            raise VariantError("allele0 is empty (length = 0)")
        if len(a1) == 0:
            raise VariantError("allele1 is empty (length = 0)")

    @classmethod
    def new_snv(cls, vcf_index, position, allele0, allele1):
        if len(allele0) != 1:
            raise VariantError("allele0 must be length 1")
        if len(allele1) != 1:
            raise VariantError("allele1 must be length 1")
        return cls(vcf_index, VariantType.Snv, position, bytes(allele0), bytes(allele1))

    @classmethod
    def new_deletion(cls, vcf_index, position, allele0, allele1):
        if len(allele0) <= 1:
            raise VariantError("reference must have length > 1")
        if len(allele1) != 1:
            raise VariantError("allele1 must be length 1")
        return cls(vcf_index, VariantType.Deletion, position, bytes(allele0), bytes(allele1))

    @classmethod
    def new_insertion(cls, vcf_index, position, allele0, allele1):
        if len(allele0) != 1:
            raise VariantError("allele0 must be length 1")
        if len(allele1) <= 1:
            raise VariantError("alternate must have length > 1")
        return cls(vcf_index, VariantType.Insertion, position, bytes(allele0), bytes(allele1))

    @classmethod
    def new_indel(cls, vcf_index, position, allele0, allele1):
        if len(allele0) <= 1:
            raise VariantError("reference must have length > 1")
        if len(allele1) <= 1:
            raise VariantError("alternate must have length > 1")
        return cls(vcf_index, VariantType.Indel, position, bytes(allele0), bytes(allele1))

    @classmethod
    def new_sv_deletion(cls, vcf_index, position, allele0, allele1):
        cls._chk(allele0, allele1)
        if len(allele1) > len(allele0):
            raise VariantError("SV deletion ALT length must be <= REF length")
        return cls(vcf_index, VariantType.SvDeletion, position, bytes(allele0), bytes(allele1))

    @classmethod
    def new_sv_insertion(cls, vcf_index, position, allele0, allele1):
        cls._chk(allele0, allele1)
        if len(allele1) < len(allele0):
            raise VariantError("SV insertion ALT length must be >= REF length")
        return cls(vcf_index, VariantType.SvInsertion, position, bytes(allele0), bytes(allele1))

    @classmethod
    def new_tr_contraction(cls, vcf_index, position, allele0, allele1):
        cls._chk(allele0, allele1)
        if len(allele1) >= len(allele0):
            raise VariantError("TR contraction ALT length must be < REF length")
        return cls(vcf_index, VariantType.TrContraction, position, bytes(allele0), bytes(allele1))

    @classmethod
    def new_tr_expansion(cls, vcf_index, position, allele0, allele1):
        cls._chk(allele0, allele1)
        if len(allele1) < len(allele0):
            raise VariantError("TR expansion ALT length must be >= REF length")
        return cls(vcf_index, VariantType.TrExpansion, position, bytes(allele0), bytes(allele1))

    def ref_len(self) -> int:
        return len(self.allele0)

    def set_raw_allele_space(self, raw_allele_space: int):
        if raw_allele_space < max(len(self.allele0), len(self.allele1)):
            raise VariantError("raw allele space must be >= allele0 and allele1 length")
        self.raw_allele_space = raw_allele_space


@dataclass
class Coordinates:
    chrom: str
    start: int
    end: int


@dataclass
class CompareRegion:
    region_id: int
    coordinates: Coordinates
    truth_variants: List[Variant]
    truth_zygosity: List[PhasedZygosity]
    query_variants: List[Variant]
    query_zygosity: List[PhasedZygosity]

    def __post_init__(self):
        if len(self.truth_variants) != len(self.truth_zygosity):
            raise ValueError("Number of truth variants and zygosities must be equal")
        if len(self.query_variants) != len(self.query_zygosity):
            raise ValueError("Number of query variants and zygosities must be equal")


@dataclass
class MultiRegion:
    region_id: int
    coordinates: Coordinates
    variants: List[List[Variant]]
    zygosity: List[List[PhasedZygosity]]

    def __post_init__(self):
        if len(self.variants) != len(self.zygosity):
            raise ValueError("variants and zygosity must have equal length")
        for v, z in zip(self.variants, self.zygosity):
            if len(v) != len(z):
                raise ValueError("variants and zygosity must have equal length")


@dataclass
class CompareConfig:
    enable_sequences: bool = True
    enable_exact_shortcut: bool = False
    max_branch_factor: int = 50
    # not in the reference: node expansions one optimize_gt_alleles call may use before the region ends with the stand-in for
    # the reference's 300 s bail-out (0 = the library default, 2^24)
    exact_gt_max_expansions: int = 0


@dataclass
class MergeConfig:
    no_conflict_enabled: bool = False
    majority_voting_enabled: bool = False
    conflict_selection: Optional[int] = None
    max_branch_factor: int = 50


@dataclass(frozen=True)
class SummaryMetrics:
    truth_tp: int = 0
    truth_fn: int = 0
    query_tp: int = 0
    query_fp: int = 0


@dataclass(frozen=True)
class SummaryGtMetrics:
    truth_tp: int = 0
    truth_fn: int = 0
    query_tp: int = 0
    query_fp: int = 0
    truth_fn_gt: int = 0
    query_fp_gt: int = 0


class GroupMetrics:
    """22 u64 counters, src/data_types/grouped_metrics.rs:149-161."""

    def __init__(self, row):
        self.row = [int(x) for x in row]

    def gt(self):
        return SummaryGtMetrics(*self.row[0:6])

    def hap(self):
        return SummaryMetrics(*self.row[abi.M_HAP:abi.M_HAP + 4])

    def weighted_hap(self):
        return SummaryMetrics(*self.row[abi.M_WEIGHTED_HAP:abi.M_WEIGHTED_HAP + 4])

    def basepair(self):
        return SummaryMetrics(*self.row[abi.M_BASEPAIR:abi.M_BASEPAIR + 4])

    def record_bp(self):
        return SummaryMetrics(*self.row[abi.M_RECORD_BP:abi.M_RECORD_BP + 4])

    def __eq__(self, other):
        return self.row == other.row


class GroupTypeMetrics:
    """joint + BTreeMap<VariantType, GroupMetrics>, grouped_metrics.rs:31-37."""

    def __init__(self, rows, mask):
        self._joint = GroupMetrics(rows[0])
        self._by_type = {VariantType(t): GroupMetrics(rows[1 + t])
                         for t in range(abi.N_VARIANT_TYPES) if mask & (1 << t)}

    def joint_metrics(self):
        return self._joint

    def variant_metrics(self):
        return self._by_type


class VariantSource(IntEnum):
    Truth = 0
    Query = 1


@dataclass(frozen=True)
class VariantMetrics:
    source: VariantSource
    classification: Classification
    expected_allele_count: int
    observed_allele_count: int

    @classmethod
    def new(cls, source, expected, observed):
        """src/data_types/variant_metrics.rs:43-71."""
        if expected > 2 or observed > 2:
            raise ValueError("allele counts must be in range: [0, 2]")
        if expected < observed:
            c = Classification.FalsePositive
        elif expected == observed:
            if expected == 0:
                raise ValueError("Variant metrics does not support expected and observed allele counts of 0")
            c = Classification.TruePositive
        else:
            c = Classification.FalseNegative
        return cls(source, c, expected, observed)

    @classmethod
    def toggle_source(cls, o):
        """src/data_types/variant_metrics.rs:77-101."""
        src = VariantSource.Query if o.source == VariantSource.Truth else VariantSource.Truth
        c = {Classification.FalseNegative: Classification.FalsePositive,
             Classification.FalsePositive: Classification.FalseNegative}.get(o.classification, o.classification)
        return cls(src, c, o.observed_allele_count, o.expected_allele_count)


@dataclass
class SequenceBundle:
    ref_seq: str
    truth_seq1: str
    truth_seq2: str
    query_seq1: str
    query_seq2: str


@dataclass
class CompareBenchmark:
    """src/data_types/compare_benchmark.rs:9-33."""
    region_id: int
    bm_edit_distance_h1: int
    bm_edit_distance_h2: int
    _group_metrics: GroupTypeMetrics
    _truth_variant_data: List[VariantMetrics]
    _query_variant_data: List[VariantMetrics]
    _sequence_bundle: Optional[SequenceBundle] = None
    containment_regions: Optional[List[int]] = None

    def total_ed(self):
        return self.bm_edit_distance_h1 + self.bm_edit_distance_h2

    def group_metrics(self):
        return self._group_metrics

    def truth_variant_data(self):
        return self._truth_variant_data

    def query_variant_data(self):
        return self._query_variant_data

    def sequence_bundle(self):
        return self._sequence_bundle


@dataclass(frozen=True)
class MergeClassification:
    """src/data_types/merge_benchmark.rs:7-20; `kind` is one of the abi.MERGE_* codes."""
    kind: int
    indices: tuple = ()

    Different = None  # filled below
    BasepairIdentical = None

    @classmethod
    def NoConflict(cls, indices):
        return cls(abi.MERGE_NO_CONFLICT, tuple(indices))

    @classmethod
    def MajorityAgree(cls, indices):
        return cls(abi.MERGE_MAJORITY_AGREE, tuple(indices))

    @classmethod
    def ConflictSelection(cls, index):
        return cls(abi.MERGE_CONFLICT_SELECTION, (index,))

    def simplify(self):
        return ["different", "no_conflict", "majority", "conflict_select", "identical"][self.kind]

    def __str__(self):
        return self.simplify() + "".join(f"_{i}" for i in self.indices)


MergeClassification.Different = MergeClassification(abi.MERGE_DIFFERENT)
MergeClassification.BasepairIdentical = MergeClassification(abi.MERGE_BASEPAIR_IDENTICAL)


@dataclass
class MergeBenchmark:
    region_id: int
    merge_classification: MergeClassification


class RegionError(RuntimeError):
    """The per-region anyhow::Error of the reference (src/main.rs:259-262)."""

    def __init__(self, region_id, status):
        names = ["ok", "unsupported zygosity", "no result found", "truth false positive",
                 "TP is less than basepair TP", "malformed region", "workspace exhausted",
                 "exact-GT expansion limit reached (stands for the reference's 300 second time limit)"]
        super().__init__(f"region {region_id}: {names[status] if status < len(names) else status}")
        self.region_id = region_id
        self.status = status
