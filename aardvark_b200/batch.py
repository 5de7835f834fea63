"""Flat SoA/CSR batches: the marshalling layer between region objects and the C ABI.

`RegionBatch` owns numpy arrays laid out exactly as `avk_region_batch`
(include/aardvark_b200.h); `CompareOutputs` / `MergeOutputs` own the
caller-allocated result arrays.  The reference materialises every region
before solving (src/main.rs:217-232), so one batch == one `aardvark compare`
(or `merge`) solve phase.
"""
import ctypes as C
from typing import List, Sequence

import numpy as np

from . import abi
from .types import CompareRegion, MultiRegion


class RegionBatch:
    """Arrays of an avk_region_batch.  Build with from_compare_regions / from_multi_regions
    or directly from arrays (the synthetic generator does the latter)."""

    def __init__(self, n_inputs, region_id, contig, start, end, var_off, position, variant_type,
                 zygosity, raw_allele_space, allele_off, a0_len, a1_len, allele_pool):
        self.n_inputs = int(n_inputs)
        self.region_id = np.ascontiguousarray(region_id, dtype=np.uint64)
        self.contig = np.ascontiguousarray(contig, dtype=np.uint32)
        self.start = np.ascontiguousarray(start, dtype=np.uint32)
        self.end = np.ascontiguousarray(end, dtype=np.uint32)
        self.var_off = np.ascontiguousarray(var_off, dtype=np.uint64)
        self.position = np.ascontiguousarray(position, dtype=np.uint32)
        self.variant_type = np.ascontiguousarray(variant_type, dtype=np.uint8)
        self.zygosity = np.ascontiguousarray(zygosity, dtype=np.uint8)
        self.raw_allele_space = np.ascontiguousarray(raw_allele_space, dtype=np.uint32)
        self.allele_off = np.ascontiguousarray(allele_off, dtype=np.uint32)
        self.a0_len = np.ascontiguousarray(a0_len, dtype=np.uint32)
        self.a1_len = np.ascontiguousarray(a1_len, dtype=np.uint32)
        self.allele_pool = np.ascontiguousarray(allele_pool, dtype=np.uint8)
        if self.allele_pool.size == 0:
            self.allele_pool = np.zeros(1, dtype=np.uint8)
        n = self.n_regions
        assert self.var_off.size == n * self.n_inputs + 1
        assert self.contig.size == n and self.start.size == n and self.end.size == n
        nv = self.n_variants
        for a in (self.variant_type, self.zygosity, self.raw_allele_space, self.allele_off, self.a0_len, self.a1_len):
            assert a.size == nv

    @property
    def n_regions(self):
        return int(self.region_id.size)

    @property
    def n_variants(self):
        return int(self.position.size)

    def nbytes(self):
        return sum(getattr(self, k).nbytes for k in (
            "region_id", "contig", "start", "end", "var_off", "position", "variant_type", "zygosity",
            "raw_allele_space", "allele_off", "a0_len", "a1_len", "allele_pool"))

    def to_c(self) -> abi.RegionBatch:
        vt = abi.VariantTable(
            self.n_variants, abi.ptr(self.position), abi.ptr(self.variant_type), abi.ptr(self.zygosity),
            abi.ptr(self.raw_allele_space), abi.ptr(self.allele_off), abi.ptr(self.a0_len), abi.ptr(self.a1_len),
            abi.ptr(self.allele_pool), self.allele_pool.size)
        return abi.RegionBatch(self.n_regions, self.n_inputs, abi.ptr(self.region_id), abi.ptr(self.contig),
                               abi.ptr(self.start), abi.ptr(self.end), abi.ptr(self.var_off), vt)

    # -- construction from region objects -----------------------------------
    @classmethod
    def _from_lists(cls, n_inputs, ids, contigs, starts, ends, lists):
        """lists: per region, per input, [(Variant, zyg_code)]"""
        var_off = [0]
        pos, vt, zy, raw, aoff, l0, l1 = [], [], [], [], [], [], []
        pool = bytearray()
        for region_lists in lists:
            assert len(region_lists) == n_inputs
            for lst in region_lists:
                for v, z in lst:
                    pos.append(v.position)
                    vt.append(int(v.variant_type))
                    zy.append(int(z))
                    raw.append(v.raw_allele_space)
                    aoff.append(len(pool))
                    l0.append(len(v.allele0))
                    l1.append(len(v.allele1))
                    pool += v.allele0
                    pool += v.allele1
                var_off.append(len(pos))
        return cls(n_inputs, ids, contigs, starts, ends, var_off, pos, vt, zy, raw, aoff, l0, l1,
                   np.frombuffer(bytes(pool), dtype=np.uint8))

    @classmethod
    def from_compare_regions(cls, regions: Sequence[CompareRegion], contig_index):
        """contig_index: dict chrom name -> contig number used in set_reference."""
        return cls._from_lists(
            2, [r.region_id for r in regions], [contig_index[r.coordinates.chrom] for r in regions],
            [r.coordinates.start for r in regions], [r.coordinates.end for r in regions],
            [[list(zip(r.truth_variants, r.truth_zygosity)), list(zip(r.query_variants, r.query_zygosity))]
             for r in regions])

    @classmethod
    def from_multi_regions(cls, regions: Sequence[MultiRegion], contig_index):
        k = len(regions[0].variants) if regions else 2
        return cls._from_lists(
            k, [r.region_id for r in regions], [contig_index[r.coordinates.chrom] for r in regions],
            [r.coordinates.start for r in regions], [r.coordinates.end for r in regions],
            [[list(zip(v, z)) for v, z in zip(r.variants, r.zygosity)] for r in regions])

    @staticmethod
    def concat(batches, contigs=None):
        """Batches over different contigs -> one batch (region ids renumbered in order; batch k gets contig index
        contigs[k], default k).  allele_off is u32: the joined allele pool must stay below 4 GiB."""
        k = batches[0].n_inputs
        assert all(b.n_inputs == k for b in batches)
        contigs = list(range(len(batches))) if contigs is None else list(contigs)
        v0 = np.cumsum([0] + [b.n_variants for b in batches])
        p0 = np.cumsum([0] + [int(b.allele_pool.size) if b.n_variants else 0 for b in batches])
        assert int(p0[-1]) < (1 << 32)
        var_off = np.concatenate([b.var_off[:-1] + np.uint64(v0[i]) for i, b in enumerate(batches)] + [np.array([v0[-1]], dtype=np.uint64)])
        cat = lambda name: np.concatenate([getattr(b, name) for b in batches])
        n = sum(b.n_regions for b in batches)
        return RegionBatch(
            k, np.arange(n, dtype=np.uint64), np.concatenate([np.full(b.n_regions, contigs[i], dtype=np.uint32) for i, b in enumerate(batches)]),
            cat("start"), cat("end"), var_off, cat("position"), cat("variant_type"), cat("zygosity"), cat("raw_allele_space"),
            np.concatenate([b.allele_off + np.uint32(p0[i]) for i, b in enumerate(batches)]), cat("a0_len"), cat("a1_len"),
            np.concatenate([b.allele_pool if b.n_variants else b.allele_pool[:0] for b in batches]))

    def slice_regions(self, lo, hi):
        """Contiguous region bin [lo, hi) as its own batch (multi-GPU sharding)."""
        k = self.n_inputs
        v0 = int(self.var_off[lo * k])
        v1 = int(self.var_off[hi * k])
        if v1 > v0:
            ao = self.allele_off[v0:v1].astype(np.int64)
            if ao.size > 1 and np.any(np.diff(ao) < 0):
                raise ValueError("slice_regions needs an allele pool laid out in variant order (allele_off not monotone)")
            p0 = int(self.allele_off[v0])
            p1 = int(self.allele_off[v1 - 1]) + int(self.a0_len[v1 - 1]) + int(self.a1_len[v1 - 1])
        else:
            p0 = p1 = 0
        return RegionBatch(
            k, self.region_id[lo:hi], self.contig[lo:hi], self.start[lo:hi], self.end[lo:hi],
            self.var_off[lo * k:hi * k + 1] - np.uint64(v0), self.position[v0:v1], self.variant_type[v0:v1],
            self.zygosity[v0:v1], self.raw_allele_space[v0:v1], self.allele_off[v0:v1] - np.uint32(p0),
            self.a0_len[v0:v1], self.a1_len[v0:v1], self.allele_pool[p0:p1])


class CallSets:
    """K call sets of one contig as one variant table (avk_callsets): input k = variants [input_off[k], input_off[k+1]),
    each in VCF order.  Records are (position, allele0, allele1, zygosity code, type code, raw_allele_space)."""

    def __init__(self, inputs, contigs=None):
        """contigs: per input, the contig index of every record (several contigs in one table, avk_build_regions_bed)."""
        pos, vt, zy, raw, aoff, l0, l1 = [], [], [], [], [], [], []
        pool = bytearray()
        off = [0]
        self.variant_contig = None if contigs is None else np.asarray([c for lst in contigs for c in lst], dtype=np.uint32)
        for lst in inputs:
            for (p_, a0, a1, z, t, rw) in lst:
                pos.append(p_); vt.append(t); zy.append(z); raw.append(rw)
                aoff.append(len(pool)); l0.append(len(a0)); l1.append(len(a1))
                pool.extend(a0); pool.extend(a1)
            off.append(len(pos))
        self.n_inputs = len(inputs)
        self.input_off = np.asarray(off, dtype=np.uint64)
        self.position = np.asarray(pos, dtype=np.uint32)
        self.variant_type = np.asarray(vt, dtype=np.uint8)
        self.zygosity = np.asarray(zy, dtype=np.uint8)
        self.raw_allele_space = np.asarray(raw, dtype=np.uint32)
        self.allele_off = np.asarray(aoff, dtype=np.uint32)
        self.a0_len = np.asarray(l0, dtype=np.uint32)
        self.a1_len = np.asarray(l1, dtype=np.uint32)
        self.allele_pool = np.frombuffer(bytes(pool) or b"\0", dtype=np.uint8).copy()
        self.pool_len = len(pool)

    @property
    def n_variants(self):
        return int(self.position.size)

    def to_c(self) -> abi.CallSets:
        vt = abi.VariantTable(
            self.n_variants, abi.ptr(self.position), abi.ptr(self.variant_type), abi.ptr(self.zygosity),
            abi.ptr(self.raw_allele_space), abi.ptr(self.allele_off), abi.ptr(self.a0_len), abi.ptr(self.a1_len),
            abi.ptr(self.allele_pool), self.pool_len)
        return abi.CallSets(self.n_inputs, abi.ptr(self.input_off), vt)


class BedIntervals:
    """High-confidence intervals per contig (avk_bed_intervals): per_contig[c] = [(start, end), ...], 0-based half-open, sorted,
    non-overlapping; one list per contig of the reference, in reference order."""

    def __init__(self, per_contig):
        self.n_contigs = len(per_contig)
        self.first = np.asarray(np.cumsum([0] + [len(x) for x in per_contig]), dtype=np.uint64)
        flat = [iv for lst in per_contig for iv in lst]
        self.start = np.asarray([a for a, _ in flat] or [0], dtype=np.uint32)
        self.end = np.asarray([b for _, b in flat] or [0], dtype=np.uint32)

    def to_c(self) -> abi.BedIntervals:
        return abi.BedIntervals(self.n_contigs, abi.ptr(self.first), abi.ptr(self.start), abi.ptr(self.end))


class CompareOutputs:
    """Caller-allocated arrays of an avk_compare_out."""

    def __init__(self, batch: RegionBatch, region_metrics=True, strat_off=None, strat_idx=None, n_strata=0,
                 seq_off=None, seq_pool_len=0, containment=False, device_strata=False):
        """containment: ask for the per-region containment masks of the device lookup (Solver.set_stratifications);
        device_strata: stratified sums over that lookup instead of a strat_off / strat_idx membership list."""
        n, nv = batch.n_regions, batch.n_variants
        self.status = np.full(max(n, 1), -1, dtype=np.int32)
        self.ed1 = np.zeros(max(n, 1), dtype=np.uint32)
        self.ed2 = np.zeros(max(n, 1), dtype=np.uint32)
        self.region_metrics = (np.zeros((max(n, 1), abi.N_GROUPS, abi.N_METRICS), dtype=np.uint64)
                               if region_metrics else None)
        self.type_mask = np.zeros(max(n, 1), dtype=np.uint16)
        self.var_expected = np.zeros(max(nv, 1), dtype=np.uint8)
        self.var_observed = np.zeros(max(nv, 1), dtype=np.uint8)
        self.var_class = np.zeros(max(nv, 1), dtype=np.uint8)
        self.totals = np.zeros((abi.N_GROUPS, abi.N_METRICS), dtype=np.uint64)
        self.totals_mask = np.zeros(1, dtype=np.uint16)
        self.solved_blocks = np.zeros(1, dtype=np.uint64)
        self.error_blocks = np.zeros(1, dtype=np.uint64)
        self.strat_off = None if strat_off is None else np.ascontiguousarray(strat_off, dtype=np.uint64)
        self.strat_idx = None if strat_idx is None else np.ascontiguousarray(strat_idx, dtype=np.uint32)
        if self.strat_idx is not None and self.strat_idx.size == 0:
            self.strat_idx = np.zeros(1, dtype=np.uint32)
        self.n_strata = int(n_strata)
        self.strat_totals = (np.zeros((max(n_strata, 1), abi.N_GROUPS, abi.N_METRICS), dtype=np.uint64)
                             if (strat_off is not None or device_strata) else None)
        self.containment = np.zeros(max(n, 1), dtype=np.uint64) if containment else None
        self.seq_off = None if seq_off is None else np.ascontiguousarray(seq_off, dtype=np.uint64)
        self.seq_len = np.zeros(max(n, 1) * 5, dtype=np.uint32) if seq_off is not None else None
        self.seq_pool = np.zeros(max(int(seq_pool_len), 1), dtype=np.uint8) if seq_off is not None else None
        self.n = n
        self.nv = nv

    def to_c(self) -> abi.CompareOut:
        return abi.CompareOut(
            abi.ptr(self.status), abi.ptr(self.ed1), abi.ptr(self.ed2), abi.ptr(self.region_metrics),
            abi.ptr(self.type_mask), abi.ptr(self.var_expected), abi.ptr(self.var_observed), abi.ptr(self.var_class),
            abi.ptr(self.totals), abi.ptr(self.totals_mask), abi.ptr(self.solved_blocks), abi.ptr(self.error_blocks),
            abi.ptr(self.strat_off), abi.ptr(self.strat_idx), self.n_strata, 0, abi.ptr(self.strat_totals),
            abi.ptr(self.seq_off), abi.ptr(self.seq_len), abi.ptr(self.seq_pool), abi.ptr(self.containment))

    FIELDS = ("status", "ed1", "ed2", "region_metrics", "type_mask", "var_expected", "var_observed", "var_class",
              "totals", "totals_mask", "solved_blocks", "error_blocks", "strat_totals", "seq_len", "containment")

    def diff(self, other) -> List[str]:
        """Names of output arrays that differ from `other` (bit-exact comparison)."""
        bad = []
        for f in self.FIELDS:
            a, b = getattr(self, f), getattr(other, f)
            if a is None or b is None:
                continue
            if not np.array_equal(a, b):
                bad.append(f)
        if self.seq_pool is not None and other.seq_pool is not None and "seq_len" not in bad:
            for i in range(self.n * 5):
                o, ln = int(self.seq_off[i]), int(self.seq_len[i])
                if not np.array_equal(self.seq_pool[o:o + ln], other.seq_pool[o:o + ln]):
                    bad.append("seq_pool")
                    break
        return bad

    def sequence(self, r, s) -> bytes:
        o, ln = int(self.seq_off[r * 5 + s]), int(self.seq_len[r * 5 + s])
        return self.seq_pool[o:o + ln].tobytes()


class MergeOutputs:
    def __init__(self, batch: RegionBatch):
        n, k = batch.n_regions, batch.n_inputs
        self.status = np.full(max(n, 1), -1, dtype=np.int32)
        self.classification = np.zeros(max(n, 1), dtype=np.uint8)
        self.n_indices = np.zeros(max(n, 1), dtype=np.uint8)
        self.indices = np.full((max(n, 1), k), 0xFF, dtype=np.uint8)

    def to_c(self) -> abi.MergeOut:
        return abi.MergeOut(abi.ptr(self.status), abi.ptr(self.classification), abi.ptr(self.n_indices),
                            abi.ptr(self.indices))

    def diff(self, other) -> List[str]:
        return [f for f in ("status", "classification", "n_indices", "indices")
                if not np.array_equal(getattr(self, f), getattr(other, f))]


def seq_offsets(batch: RegionBatch):
    """Upper-bound layout for the optional sequence bundle (mirrors avk_compare_seq_offsets):
    ref = window length; each haplotype <= window + sum of ALT lengths of its side."""
    n, k = batch.n_regions, batch.n_inputs
    assert k == 2
    win = (batch.end.astype(np.int64) - batch.start.astype(np.int64)).clip(min=0)
    csum = np.concatenate([[0], np.cumsum(batch.a1_len.astype(np.int64))])
    vo = batch.var_off.astype(np.int64)
    t_alt = csum[vo[1::2]] - csum[vo[0:-1:2]]
    q_alt = csum[vo[2::2]] - csum[vo[1::2]]
    sizes = np.stack([win, win + t_alt, win + t_alt, win + q_alt, win + q_alt], axis=1).reshape(-1)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    return off, int(off[-1])


class StratIntervals:
    """Stratification interval sets for the device containment lookup (avk_strat_intervals): `strata` is a list (one entry
    per stratum, in label order) of lists of (contig index, BED start, BED end) -- half-open BED coordinates, stored 0-based
    inclusive as the reference does (stratifications.rs:158-163)."""

    def __init__(self, strata, n_contigs):
        self.n_strata, self.n_contigs = len(strata), int(n_contigs)
        off, first, last = [0], [], []
        for ivs in strata:
            for c in range(self.n_contigs):
                for (cc, s, e) in ivs:
                    if cc == c:
                        first.append(s)
                        last.append(e - 1)
                off.append(len(first))
        self.off = np.asarray(off, dtype=np.uint64)
        self.first = np.asarray(first if first else [0], dtype=np.uint32)
        self.last = np.asarray(last if last else [0], dtype=np.uint32)

    def to_c(self) -> abi.StratIntervals:
        return abi.StratIntervals(self.n_strata, self.n_contigs, abi.ptr(self.off), abi.ptr(self.first), abi.ptr(self.last))


def masks_to_membership(masks: np.ndarray):
    """containment bit masks -> (strat_off, strat_idx) membership lists"""
    off, idx = [0], []
    for m in masks.tolist():
        s = 0
        while m:
            if m & 1:
                idx.append(s)
            m >>= 1
            s += 1
        off.append(len(idx))
    return np.asarray(off, dtype=np.uint64), np.asarray(idx if idx else [0], dtype=np.uint32)
