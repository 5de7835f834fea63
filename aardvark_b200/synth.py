"""Synthetic workloads of the shapes BASELINE.json names (there is no network for real data).

A random reference with low-complexity inserts, a truth call set, and query call
sets derived from it with injected representation differences (SNV pairs <-> MNP
records, tandem-duplication insertions / repeat deletions shifted inside their
repeat), genotype errors, FNs and FPs.  Variants are trimmed and typed as
src/parsing/region_generation.rs:612-626, 719-758 and clustered exactly as
src/parsing/region_generation.rs:396-469: one BED interval
per contig, `flank` = --min-variant-gap.

Everything is seeded; the same (config, seed) yields the same batch on every box.
"""
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from . import abi
from .batch import RegionBatch

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)

# a variant record while generating: (pos, allele0 bytes, allele1 bytes, zyg code, type code, raw_allele_space)
Rec = Tuple[int, bytes, bytes, int, int, int]


def random_reference(length: int, rng: np.random.Generator, low_complexity_frac: float = 0.02) -> np.ndarray:
    """i.i.d. uniform ACGT plus homopolymers (8-30 bp) and STRs (unit 2-6 x 5-40) covering about
    `low_complexity_frac` of the contig, so that indel shift ambiguity actually occurs."""
    ref = ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]
    if length < 200 or low_complexity_frac <= 0:
        return ref
    n_rep = max(1, int(length * low_complexity_frac / 40))
    starts = rng.integers(0, max(1, length - 260), size=n_rep)
    kinds = rng.random(n_rep)
    for s, k in zip(starts.tolist(), kinds.tolist()):
        if k < 0.5:
            ln = int(rng.integers(8, 31))
            ref[s:s + ln] = ACGT[rng.integers(0, 4)]
        else:
            unit = ACGT[rng.integers(0, 4, size=int(rng.integers(2, 7)))]
            reps = int(rng.integers(5, 41))
            blk = np.tile(unit, reps)[:min(unit.size * reps, length - s)]
            ref[s:s + blk.size] = blk
    return ref


def infer_type(a0: bytes, a1: bytes, sv: bool = False) -> int:
    """Type by allele lengths (region_generation.rs:749-755) or SVTYPE-style flag (:723-737)."""
    if sv:
        return abi.VT_SV_INSERTION if len(a1) >= len(a0) else abi.VT_SV_DELETION
    l0, l1 = len(a0), len(a1)
    if l0 == 1 and l1 == 1:
        return abi.VT_SNV
    if l0 == 1:
        return abi.VT_INSERTION
    if l1 == 1:
        return abi.VT_DELETION
    return abi.VT_INDEL


def make_rec(pos: int, a0: bytes, a1: bytes, zyg: int, sv: bool = False) -> Rec:
    """raw_allele_space before trimming (:612); trailing identical bases trimmed while both
    alleles are longer than one base (:615-618)."""
    raw = max(len(a0), len(a1))
    while len(a0) > 1 and len(a1) > 1 and a0[-1] == a1[-1]:
        a0, a1 = a0[:-1], a1[:-1]
    return (pos, a0, a1, zyg, infer_type(a0, a1, sv), raw)


@dataclass
class SynthParams:
    n_variants: int = 150_000          # truth variants on the contig
    snv_frac: float = 0.85
    indel_geom_mean: float = 3.0
    indel_cap: int = 49
    dense_frac: float = 0.15           # fraction of spacings drawn from the dense component
    dense_mean: float = 12.0
    het_frac: float = 0.62
    phased_frac: float = 0.70
    # query derivation (per truth variant)
    p_repr: float = 0.02
    p_gt_err: float = 0.015
    p_fn: float = 0.015
    p_fp: float = 0.01
    flank: int = 50                    # --min-variant-gap
    # SV / long-indel component (config 4)
    sv_events: int = 0
    sv_min: int = 50
    sv_max: int = 10_000
    sv_p_shift: float = 0.10
    sv_p_diverge: float = 0.10
    sv_p_fn: float = 0.05
    sv_p_fp: float = 0.05


def _rand_seq(rng, n) -> bytes:
    return ACGT[rng.integers(0, 4, size=n)].tobytes()


def _rand_zyg(rng, p: SynthParams) -> int:
    if rng.random() < p.het_frac:
        if rng.random() < p.phased_frac:
            return abi.ZYG_PHASED_HET01 if rng.random() < 0.5 else abi.ZYG_PHASED_HET10
        return abi.ZYG_UNPHASED_HET
    return abi.ZYG_HOM_ALT


def gen_truth(ref: np.ndarray, p: SynthParams, rng: np.random.Generator) -> List[Rec]:
    """Poisson-spaced small variants with real-like clustering, plus optional SV events."""
    L = ref.size
    n = p.n_variants
    recs: List[Rec] = []
    if n > 0:
        sparse_mean = max(2.0, (L / n - p.dense_frac * p.dense_mean) / (1.0 - p.dense_frac))
        dense = rng.random(n) < p.dense_frac
        gaps = np.where(dense, rng.exponential(p.dense_mean, n), rng.exponential(sparse_mean, n))
        pos = np.cumsum(np.maximum(1, gaps.astype(np.int64))) + p.flank
        pos = pos[pos < L - p.indel_cap - p.flank - 2]
        kinds = rng.random(pos.size)
        lens = np.minimum(p.indel_cap, rng.geometric(1.0 / p.indel_geom_mean, pos.size))
        last_end = -1
        for i in range(pos.size):
            q = int(pos[i])
            if q <= last_end and rng.random() < 0.9:   # keep a few overlapping records (conflict paths)
                continue
            z = _rand_zyg(rng, p)
            k = kinds[i]
            if k < p.snv_frac:
                r = int(ref[q])
                alt = int(ACGT[(np.searchsorted(ACGT, r) + 1 + rng.integers(0, 3)) % 4])
                rec = make_rec(q, bytes([r]), bytes([alt]), z)
            else:
                ln = int(lens[i])
                anchor = bytes([int(ref[q])])
                if rng.random() < 0.5:        # insertion; half of them tandem duplications of the following bases
                    ins = ref[q + 1:q + 1 + ln].tobytes() if rng.random() < 0.5 else _rand_seq(rng, ln)
                    rec = make_rec(q, anchor, anchor + ins, z)
                else:
                    rec = make_rec(q, ref[q:q + 1 + ln].tobytes(), anchor, z)
            recs.append(rec)
            last_end = max(last_end, q + len(rec[1]) - 1)
    if p.sv_events > 0:
        sv_pos = np.sort(rng.integers(p.flank + 1, L - p.sv_max - p.flank - 2, size=p.sv_events))
        sv_len = np.exp(rng.uniform(np.log(p.sv_min), np.log(p.sv_max), size=p.sv_events)).astype(np.int64)
        last_end = -1
        for q, ln in zip(sv_pos.tolist(), sv_len.tolist()):
            if q <= last_end:
                continue
            z = _rand_zyg(rng, p)
            anchor = bytes([int(ref[q])])
            if rng.random() < 0.5:
                ins = ref[q + 1:q + 1 + ln].tobytes() if rng.random() < 0.3 else _rand_seq(rng, ln)
                rec = make_rec(q, anchor, anchor + ins, z, sv=True)
            else:
                rec = make_rec(q, ref[q:q + 1 + ln].tobytes(), anchor, z, sv=True)
            recs.append(rec)
            last_end = q + len(rec[1]) - 1
        recs.sort(key=lambda r: r[0])
    return recs


def _requery_zyg(z: int, rng) -> int:
    """Query phase is randomised: the solver must recover the orientation."""
    if z == abi.ZYG_HOM_ALT:
        return z
    u = rng.random()
    return abi.ZYG_UNPHASED_HET if u < 0.5 else (abi.ZYG_PHASED_HET01 if u < 0.75 else abi.ZYG_PHASED_HET10)


def _shift_insertion(ref, pos, a0, a1, rng):
    """Right-shift a tandem-duplication insertion inside its repeat: same haplotype, other record."""
    ins = a1[1:]
    L = len(ins)
    if len(a0) != 1 or L == 0 or pos + 1 + L >= ref.size or ref[pos + 1:pos + 1 + L].tobytes() != ins:
        return None
    s = int(rng.integers(1, L + 1))
    anchor = bytes([int(ref[pos + s])])
    return pos + s, anchor, anchor + ins[s:] + ins[:s]


def _shift_deletion(ref, pos, a0, a1, rng):
    """Right-shift a deletion while the base entering equals the base leaving."""
    L = len(a0) - 1
    if len(a1) != 1 or L <= 0:
        return None
    s = 0
    while s < 2 * L and pos + 1 + s + L < ref.size and ref[pos + 1 + s] == ref[pos + 1 + s + L]:
        s += 1
    if s == 0:
        return None
    s = int(rng.integers(1, s + 1))
    return pos + s, ref[pos + s:pos + s + 1 + L].tobytes(), bytes([int(ref[pos + s])])


def derive_query(ref: np.ndarray, truth: Sequence[Rec], p: SynthParams, rng: np.random.Generator,
                 err_scale: float = 1.0, dropout_blocks: float = 0.0) -> List[Rec]:
    """Query call set = truth transformed record by record (SURVEY.md 8d generator)."""
    out: List[Rec] = []
    n = len(truth)
    u = rng.random(n)
    i = 0
    p_repr, p_gt, p_fn, p_fp = (x * err_scale for x in (p.p_repr, p.p_gt_err, p.p_fn, p.p_fp))
    while i < n:
        pos, a0, a1, z, vt, raw = truth[i]
        is_sv = vt in (abi.VT_SV_INSERTION, abi.VT_SV_DELETION)
        x = u[i]
        if is_sv:
            zq = _requery_zyg(z, rng)
            if x < p.sv_p_fn:
                pass
            elif x < p.sv_p_fn + p.sv_p_shift:
                sh = _shift_insertion(ref, pos, a0, a1, rng) if vt == abi.VT_SV_INSERTION else _shift_deletion(ref, pos, a0, a1, rng)
                if sh is None:
                    out.append((pos, a0, a1, zq, vt, raw))
                else:
                    out.append(make_rec(sh[0], sh[1], sh[2], zq, sv=True))
            elif x < p.sv_p_fn + p.sv_p_shift + p.sv_p_diverge and vt == abi.VT_SV_INSERTION:
                arr = np.frombuffer(a1, dtype=np.uint8).copy()
                k = max(1, int(arr.size * rng.uniform(0.01, 0.05)))
                idx = rng.integers(1, arr.size, size=k)
                arr[idx] = ACGT[rng.integers(0, 4, size=k)]
                out.append(make_rec(pos, a0, arr.tobytes(), zq, sv=True))
            else:
                out.append((pos, a0, a1, zq, vt, raw))
            if rng.random() < p.sv_p_fp:
                q = pos + int(rng.integers(20, 400))
                if q + 300 < ref.size:
                    anchor = bytes([int(ref[q])])
                    out.append(make_rec(q, anchor, anchor + _rand_seq(rng, int(rng.integers(50, 300))), _rand_zyg(rng, p), sv=True))
            i += 1
            continue
        if x < p_fn:                                     # false negative: dropped
            i += 1
            continue
        zq = _requery_zyg(z, rng)
        if x < p_fn + p_gt:                              # genotype error het <-> hom
            zq = abi.ZYG_HOM_ALT if z != abi.ZYG_HOM_ALT else abi.ZYG_UNPHASED_HET
            out.append((pos, a0, a1, zq, vt, raw))
        elif x < p_fn + p_gt + p_repr:                   # representation change
            done = False
            adv = 1
            if vt == abi.VT_SNV and i + 1 < n:
                npos, na0, na1, nz, nvt, _ = truth[i + 1]
                same_phase = (z == nz) and z != abi.ZYG_UNPHASED_HET
                if nvt == abi.VT_SNV and 0 < npos - pos <= 3 and same_phase:
                    # two SNVs on one haplotype -> one MNP/indel record (waffle_solver.rs:932-978)
                    mid = ref[pos + 1:npos].tobytes()
                    out.append(make_rec(pos, a0 + mid + na0, a1 + mid + na1, zq))
                    adv = 2
                    done = True
            if not done and vt == abi.VT_SNV and pos > 0:
                # SNV written as a 2-base record with a shared leading base
                lead = bytes([int(ref[pos - 1])])
                out.append(make_rec(pos - 1, lead + a0, lead + a1, zq))
                done = True
            if not done and vt == abi.VT_INSERTION:
                sh = _shift_insertion(ref, pos, a0, a1, rng)
                if sh is not None:
                    out.append(make_rec(sh[0], sh[1], sh[2], zq))
                    done = True
            if not done and vt == abi.VT_DELETION:
                sh = _shift_deletion(ref, pos, a0, a1, rng)
                if sh is not None:
                    out.append(make_rec(sh[0], sh[1], sh[2], zq))
                    done = True
            if not done:
                out.append((pos, a0, a1, zq, vt, raw))
            i += adv
            continue
        else:
            out.append((pos, a0, a1, zq, vt, raw))
        if rng.random() < p_fp:                          # false positive near a real variant
            q = pos + int(rng.integers(1, 40))
            if q + 2 < ref.size:
                r = int(ref[q])
                alt = int(ACGT[(np.searchsorted(ACGT, r) + 1 + rng.integers(0, 3)) % 4])
                out.append(make_rec(q, bytes([r]), bytes([alt]), _rand_zyg(rng, p)))
        i += 1
    out.sort(key=lambda r: r[0])    # stable: equal positions keep generation order
    if dropout_blocks > 0:          # per-set dropout of whole neighbourhoods (merge config)
        keep = []
        blk = -1
        drop = False
        for r in out:
            b = r[0] // 2000
            if b != blk:
                blk = b
                drop = rng.random() < dropout_blocks
            if not drop:
                keep.append(r)
        out = keep
    return out


def cluster_regions(inputs: Sequence[Sequence[Rec]], contig_len: int, flank: int, contig: int = 0,
                    first_region_id: int = 0) -> RegionBatch:
    """Gap-based clustering of K call sets, src/parsing/region_generation.rs:352-469 with ONE BED
    interval spanning the contig: inputs concatenated in input order, stable-sorted by position;
    window_start = first pos - flank (saturating); window_end = max(pos + ref_len + flank) clipped to
    the contig; a variant with pos >= window_end starts the next region."""
    K = len(inputs)
    allv = [(r[0], k, r) for k, lst in enumerate(inputs) for r in lst]
    allv.sort(key=lambda t: t[0])   # stable; ties keep input order then list order
    region_id, contigs, starts, ends = [], [], [], []
    var_off = [0]
    pos, vt, zy, raw, aoff, l0, l1 = [], [], [], [], [], [], []
    pool = bytearray()

    cur: List[List[Rec]] = [[] for _ in range(K)]
    w_start = None
    w_end = None

    def flush():
        nonlocal cur
        region_id.append(first_region_id + len(region_id))
        contigs.append(contig)
        starts.append(w_start)
        ends.append(w_end)
        for lst in cur:
            for (p_, a0, a1, z, t, rw) in lst:
                pos.append(p_)
                vt.append(t)
                zy.append(z)
                raw.append(rw)
                aoff.append(len(pool))
                l0.append(len(a0))
                l1.append(len(a1))
                pool.extend(a0)
                pool.extend(a1)
            var_off.append(len(pos))
        cur = [[] for _ in range(K)]

    for p_, k, rec in allv:
        if p_ + len(rec[1]) > contig_len:
            continue                                    # not fully contained in the BED span (:551)
        if w_end is not None and p_ >= w_end:
            flush()
            w_start = None
        if w_start is None:
            w_start = max(0, p_ - flank)
        vend = min(p_ + len(rec[1]) + flank, contig_len)
        w_end = vend if w_end is None else max(w_end, vend)
        cur[k].append(rec)
    if w_start is not None and any(cur):
        flush()
    return RegionBatch(K, region_id, contigs, starts, ends, var_off, pos, vt, zy, raw, aoff, l0, l1,
                       np.frombuffer(bytes(pool), dtype=np.uint8))


# ----------------------------------------------------------------------------- named workloads
CHR20_LEN = 64_444_167


def callsets_compare(contig_len: int, params: SynthParams, seed: int):
    """(reference contig, [truth records, query records]): the call sets before clustering."""
    rng = np.random.default_rng(seed)
    ref = random_reference(contig_len, rng)
    truth = gen_truth(ref, params, rng)
    query = derive_query(ref, truth, params, rng)
    return ref, [truth, query]


def workload_compare(contig_len: int, params: SynthParams, seed: int):
    """(reference contig as uint8 array, compare RegionBatch with n_inputs == 2)."""
    ref, inputs = callsets_compare(contig_len, params, seed)
    return ref, cluster_regions(inputs, contig_len, params.flank)


def workload_chr20(scale: float = 1.0, seed: int = 20):
    """BASELINE.json configs[1]: chr20-scale SNV + small-indel compare (~150k variants per side).
    `scale` shrinks contig and variant count together (same density) for tests."""
    L = max(2000, int(CHR20_LEN * scale))
    return workload_compare(L, SynthParams(n_variants=max(1, int(150_000 * scale))), seed)


GRCH38_LENS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422,
               135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167,
               46709983, 50818468, 156040895, 57227415]


def _wgs_contig(args):
    i, length, n_variants, seed = args
    ref, batch = workload_compare(length, SynthParams(n_variants=n_variants), seed)
    return i, ref, batch


def workload_wgs(scale: float = 1.0, seed: int = 38, workers: int = 1):
    """BASELINE.json configs[2]: whole-genome HG002-like compare -- 24 contigs with GRCh38 lengths (3.09 Gbp at scale 1),
    ~5 M variants per side, ~3.95 M clusters; contig i uses seed + i.  `scale` shrinks contig lengths and variant counts
    together.  Returns ([contig arrays], one multi-contig RegionBatch in region_id order).  `workers` > 1 generates the
    contigs in forked worker processes -- call it BEFORE the process initialises CUDA."""
    jobs = []
    for i, L in enumerate(GRCH38_LENS):
        Ls = max(20000, int(L * scale))
        jobs.append((i, Ls, max(1, int(5_000_000 * Ls / 3.1e9)), seed + i))
    if workers > 1:
        import multiprocessing as mp
        order = sorted(jobs, key=lambda j: -j[1])          # longest contigs first
        with mp.get_context("fork").Pool(min(workers, len(jobs))) as pool:
            parts = sorted(pool.imap_unordered(_wgs_contig, order), key=lambda t: t[0])
    else:
        parts = [_wgs_contig(j) for j in jobs]
    return [p[1] for p in parts], RegionBatch.concat([p[2] for p in parts])


def workload_sv(scale: float = 1.0, seed: int = 4):
    """BASELINE.json configs[3]: SV / long-indel heavy compare (50 bp - 10 kbp events, min gap 1000),
    with background small variants."""
    L = max(200_000, int(250_000_000 * scale))
    p = SynthParams(n_variants=max(1, int(390_000 * scale)), sv_events=max(1, int(20_000 * scale)), flank=1000)
    return workload_compare(L, p, seed)


def workload_merge(contig_len: int, n_variants: int, n_sets: int = 5, seed: int = 38,
                   err_scales=(0.25, 0.5, 0.5, 1.0, 1.5), dropout: float = 0.05):
    """BASELINE.json configs[4]: K call sets derived independently from one truth."""
    ref, sets, flank = callsets_merge(contig_len, n_variants, n_sets, seed, err_scales, dropout)
    return ref, cluster_regions(sets, contig_len, flank)


def callsets_merge(contig_len: int, n_variants: int, n_sets: int = 5, seed: int = 38,
                   err_scales=(0.25, 0.5, 0.5, 1.0, 1.5), dropout: float = 0.05):
    """(reference contig, K call sets, flank): the merge inputs before clustering."""
    rng = np.random.default_rng(seed)
    ref = random_reference(contig_len, rng)
    p = SynthParams(n_variants=n_variants)
    truth = gen_truth(ref, p, rng)
    sets = [derive_query(ref, truth, p, rng, err_scale=err_scales[k % len(err_scales)], dropout_blocks=dropout)
            for k in range(n_sets)]
    return ref, sets, p.flank
