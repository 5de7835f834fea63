"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs only; the product package aardvark_b200 never imports it.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from aardvark_b200 import abi  # noqa: E402
from aardvark_b200.batch import CompareOutputs, MergeOutputs, RegionBatch  # noqa: E402

_LIB = None
U64MAX = (1 << 64) - 1


def _cpu_tag():
    """Identifies the host CPU: the oracle is built with -march=native, so a library built on another
    machine (it travels with the repo snapshot) is rebuilt before it is loaded."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import hashlib
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    tag_file = so + ".cpu"
    srcs = [os.path.join(_HERE, "aardvark_oracle.cpp"), os.path.join(_HERE, "oracle.h"), os.path.join(_HERE, "Makefile"),
            os.path.join(_HERE, "..", "include", "aardvark_b200.h")]
    tag = _cpu_tag()
    try:
        built_for = open(tag_file).read().strip()
    except OSError:
        built_for = ""
    if force or not os.path.exists(so) or built_for != tag or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-B", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
        with open(tag_file, "w") as f:
            f.write(tag)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        u8p, u64 = C.POINTER(C.c_uint8), C.c_uint64
        _LIB.orc_wfa_ed.restype = u64
        _LIB.orc_wfa_ed.argtypes = [C.c_char_p, u64, C.c_char_p, u64]
        _LIB.orc_edit_distance.restype = u64
        _LIB.orc_edit_distance.argtypes = [C.c_char_p, u64, C.c_char_p, u64]
        _LIB.orc_dwfa_new.restype = C.c_void_p
        _LIB.orc_dwfa_new.argtypes = [u64]
        _LIB.orc_dwfa_clone.restype = C.c_void_p
        _LIB.orc_dwfa_clone.argtypes = [C.c_void_p]
        _LIB.orc_dwfa_free.argtypes = [C.c_void_p]
        _LIB.orc_dwfa_update.argtypes = [C.c_void_p, C.c_char_p, u64, C.c_char_p, u64]
        _LIB.orc_dwfa_finalize.argtypes = [C.c_void_p, C.c_char_p, u64, C.c_char_p, u64]
        _LIB.orc_dwfa_ed.restype = u64
        _LIB.orc_dwfa_ed.argtypes = [C.c_void_p]
        _LIB.orc_dwfa_wavefront.restype = u64
        _LIB.orc_dwfa_wavefront.argtypes = [C.c_void_p, C.POINTER(u64), u64]
        _LIB.orc_dwfa_equal.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.orc_hap_new.restype = C.c_void_p
        _LIB.orc_hap_new.argtypes = [u64, u64]
        _LIB.orc_hap_free.argtypes = [C.c_void_p]
        _LIB.orc_hap_extend.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_uint32, C.c_char_p, C.c_uint32,
                                        C.c_char_p, C.c_uint32, C.c_int, C.c_int64]
        _LIB.orc_hap_finalize.argtypes = [C.c_void_p, C.c_char_p, u64]
        for f in ("orc_hap_ed", "orc_hap_skip", "orc_hap_cost"):
            getattr(_LIB, f).restype = u64
            getattr(_LIB, f).argtypes = [C.c_void_p]
        _LIB.orc_hap_seq.restype = u64
        _LIB.orc_hap_seq.argtypes = [C.c_void_p, C.c_int, u8p, u64]
        _LIB.orc_hap_alleles.restype = u64
        _LIB.orc_hap_alleles.argtypes = [C.c_void_p, C.c_int, u8p, u64]
        _LIB.orc_perform_basepair_compare.argtypes = [C.c_char_p, u64, C.c_char_p, u64, C.c_char_p, u64, C.POINTER(u64)]
        _LIB.orc_variant_delta_length.restype = C.c_int64
        _LIB.orc_variant_delta_length.argtypes = [C.POINTER(abi.RegionBatch), C.c_uint32]
    return _LIB


def wfa_ed(a: bytes, b: bytes) -> int:
    return int(lib().orc_wfa_ed(a, len(a), b, len(b)))


def edit_distance(a: bytes, b: bytes) -> int:
    return int(lib().orc_edit_distance(a, len(a), b, len(b)))


class DWFALite:
    """src/dwfa/dynamic_wfa.rs DWFALite; update/finalize return 0 ok, 1 MaxEditDistance, 2 AlreadyFinalized."""

    def __init__(self, max_edit_distance=U64MAX, _h=None):
        self._h = _h if _h is not None else lib().orc_dwfa_new(max_edit_distance)

    def __del__(self):
        if lib is not None and self._h:
            lib().orc_dwfa_free(self._h)
            self._h = None

    def clone(self):
        return DWFALite(_h=lib().orc_dwfa_clone(self._h))

    def update(self, baseline: bytes, other: bytes) -> int:
        return lib().orc_dwfa_update(self._h, baseline, len(baseline), other, len(other))

    def finalize(self, baseline: bytes, other: bytes) -> int:
        return lib().orc_dwfa_finalize(self._h, baseline, len(baseline), other, len(other))

    def edit_distance(self) -> int:
        return int(lib().orc_dwfa_ed(self._h))

    def wavefront(self):
        buf = (C.c_uint64 * 65536)()
        n = lib().orc_dwfa_wavefront(self._h, buf, 65536)
        return [int(buf[i]) for i in range(min(n, 65536))]

    def __eq__(self, other):
        return bool(lib().orc_dwfa_equal(self._h, other._h))


class HaplotypeDWFA:
    """src/dwfa/haplotype_dwfa.rs HaplotypeDWFA."""

    def __init__(self, region_start, max_edit_distance=U64MAX):
        self._h = lib().orc_hap_new(region_start, max_edit_distance)

    def __del__(self):
        if lib is not None and self._h:
            lib().orc_hap_free(self._h)
            self._h = None

    def extend_variant(self, reference: bytes, is_truth, variant, allele, sync_extension=None):
        return lib().orc_hap_extend(self._h, reference, int(is_truth), variant.position, variant.allele0,
                                    len(variant.allele0), variant.allele1, len(variant.allele1), int(allele),
                                    -1 if sync_extension is None else sync_extension)

    def finalize_dwfa(self, reference: bytes, region_end):
        return lib().orc_hap_finalize(self._h, reference, region_end)

    def edit_distance(self):
        return int(lib().orc_hap_ed(self._h))

    def total_variant_skip_distance(self):
        return int(lib().orc_hap_skip(self._h))

    def total_cost(self):
        return int(lib().orc_hap_cost(self._h))

    def sequence(self, is_truth) -> bytes:
        buf = (C.c_uint8 * 65536)()
        n = lib().orc_hap_seq(self._h, int(is_truth), buf, 65536)
        return bytes(buf[:n])

    def alleles(self, is_truth):
        buf = (C.c_uint8 * 4096)()
        n = lib().orc_hap_alleles(self._h, int(is_truth), buf, 4096)
        return list(buf[:n])


def _contig_args(contigs):
    arrs = [np.frombuffer(c, dtype=np.uint8) if isinstance(c, (bytes, bytearray)) else np.ascontiguousarray(c, dtype=np.uint8)
            for c in contigs]
    ptrs = (C.POINTER(C.c_uint8) * len(arrs))(*[a.ctypes.data_as(C.POINTER(C.c_uint8)) for a in arrs])
    lens = (C.c_uint64 * len(arrs))(*[a.size for a in arrs])
    return arrs, ptrs, lens


def compare_batch(batch: RegionBatch, contigs, cfg: abi.CompareCfg, out: CompareOutputs = None, n_threads=0,
                  work=False, **out_kwargs):
    """Run the oracle over a whole batch.  Returns CompareOutputs (and work counters if asked)."""
    if out is None:
        out = CompareOutputs(batch, **out_kwargs)
    arrs, ptrs, lens = _contig_args(contigs)
    cb, co = batch.to_c(), out.to_c()
    wc = abi.WorkCounters()
    rc = lib().orc_compare_batch(C.byref(cb), ptrs, lens, len(arrs), C.byref(cfg), C.byref(co), int(n_threads),
                                 C.byref(wc) if work else None)
    if rc != 0:
        raise RuntimeError(f"orc_compare_batch failed: {rc}")
    return (out, wc.as_dict()) if work else out


def merge_batch(batch: RegionBatch, contigs, cfg: abi.MergeCfg, n_threads=0):
    out = MergeOutputs(batch)
    arrs, ptrs, lens = _contig_args(contigs)
    cb, co = batch.to_c(), out.to_c()
    rc = lib().orc_merge_batch(C.byref(cb), ptrs, lens, len(arrs), C.byref(cfg), C.byref(co), int(n_threads), None)
    if rc != 0:
        raise RuntimeError(f"orc_merge_batch failed: {rc}")
    return out


def optimize_sequences(batch: RegionBatch, reference: bytes, max_branch_factor=50, max_results=64):
    """optimize_sequences on region 0.  Returns (status, [dict per equal-best result])."""
    cb = batch.to_c()
    nt = int(batch.var_off[1] - batch.var_off[0])
    nq = int(batch.var_off[2] - batch.var_off[1])
    nv = nt + nq
    stride = len(reference) + int(batch.a1_len.sum()) + 8
    zyg = np.zeros(max_results * max(nv, 1), dtype=np.uint8)
    num = np.zeros(max_results * 10, dtype=np.uint64)
    seqs = np.zeros(max_results * 4 * stride, dtype=np.uint8)
    n = C.c_uint32(0)
    f = lib().orc_optimize_sequences
    f.argtypes = [C.POINTER(abi.RegionBatch), C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32),
                  C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.c_uint64]
    st = f(C.byref(cb), reference, len(reference), max_branch_factor, max_results, C.byref(n), abi.ptr(zyg),
           abi.ptr(num), abi.ptr(seqs), stride)
    res = []
    for i in range(min(n.value, max_results) if st == 0 else 0):
        nn = num[i * 10:(i + 1) * 10]
        s = [seqs[(i * 4 + k) * stride:(i * 4 + k) * stride + int(nn[6 + k])].tobytes() for k in range(4)]
        res.append(dict(truth_zygosity=list(zyg[i * nv:i * nv + nt]), query_zygosity=list(zyg[i * nv + nt:(i + 1) * nv]),
                        ed1=int(nn[0]), ed2=int(nn[1]), truth_vs1=int(nn[2]), truth_vs2=int(nn[3]),
                        query_vs1=int(nn[4]), query_vs2=int(nn[5]),
                        truth_seq1=s[0], truth_seq2=s[1], query_seq1=s[2], query_seq2=s[3]))
    return st, res


def optimize_gt_alleles(batch: RegionBatch, reference: bytes):
    """optimize_gt_alleles on region 0; the batch zygosity column carries Allele codes (1 REF, 2 ALT)."""
    cb = batch.to_c()
    nt = int(batch.var_off[1] - batch.var_off[0])
    nq = int(batch.var_off[2] - batch.var_off[1])
    t = np.zeros(max(nt, 1), dtype=np.uint8)
    q = np.zeros(max(nq, 1), dtype=np.uint8)
    ne = C.c_uint64(0)
    f = lib().orc_optimize_gt_alleles
    f.argtypes = [C.POINTER(abi.RegionBatch), C.c_char_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)]
    st = f(C.byref(cb), reference, abi.ptr(t), abi.ptr(q), C.byref(ne))
    return st, list(t[:nt]), list(q[:nq]), int(ne.value)


def generate_haplotype_sequence(batch: RegionBatch, reference: bytes, hap: int):
    cb = batch.to_c()
    cap = len(reference) + int(batch.a1_len.sum()) + 8
    buf = np.zeros(cap, dtype=np.uint8)
    ln, fe = C.c_uint64(0), C.c_uint64(0)
    f = lib().orc_generate_haplotype_sequence
    f.argtypes = [C.POINTER(abi.RegionBatch), C.c_char_p, C.c_int, C.POINTER(C.c_uint8), C.c_uint64,
                  C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    st = f(C.byref(cb), reference, hap, abi.ptr(buf), cap, C.byref(ln), C.byref(fe))
    return st, buf[:ln.value].tobytes(), int(fe.value)


def perform_basepair_compare(r: bytes, t: bytes, q: bytes):
    out = (C.c_uint64 * 4)()
    lib().orc_perform_basepair_compare(r, len(r), t, len(t), q, len(q), out)
    return tuple(int(x) for x in out)


def variant_delta_length(batch: RegionBatch, k=0):
    cb = batch.to_c()
    return int(lib().orc_variant_delta_length(C.byref(cb), k))


def containments(batch: RegionBatch, strat) -> np.ndarray:
    """Stratifications::containments of every region's var_coordinates() (orc_containments); strat: batch.StratIntervals."""
    mask = np.zeros(max(batch.n_regions, 1), dtype=np.uint64)
    cb, cs = batch.to_c(), strat.to_c()
    f = lib().orc_containments
    f.argtypes = [C.POINTER(abi.RegionBatch), C.POINTER(abi.StratIntervals), C.POINTER(C.c_uint64)]
    rc = f(C.byref(cb), C.byref(cs), abi.ptr(mask))
    if rc != 0:
        raise RuntimeError(f"orc_containments failed: {rc}")
    return mask[:batch.n_regions]


def num_threads():
    return int(lib().orc_num_threads())


def build_regions(callsets, contig_len: int, flank: int, contig: int = 0, first_region_id: int = 0) -> RegionBatch:
    """Region builder restatement (region_generation.rs:352-469) -> RegionBatch."""
    nv, k = callsets.n_variants, callsets.n_inputs
    b = RegionBatch(k, np.zeros(nv, np.uint64), np.zeros(nv, np.uint32), np.zeros(nv, np.uint32), np.zeros(nv, np.uint32),
                    np.zeros(nv * k + 1, np.uint64), np.zeros(nv, np.uint32), np.zeros(nv, np.uint8), np.zeros(nv, np.uint8),
                    np.zeros(nv, np.uint32), np.zeros(nv, np.uint32), np.zeros(nv, np.uint32), np.zeros(nv, np.uint32),
                    np.zeros(max(callsets.pool_len, 1), np.uint8))
    cs, cb = callsets.to_c(), b.to_c()
    fn = lib().orc_build_regions
    fn.argtypes = [C.POINTER(abi.CallSets), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(abi.RegionBatch)]
    rc = fn(C.byref(cs), contig_len, contig, flank, first_region_id, C.byref(cb))
    if rc != 0:
        raise RuntimeError(f"orc_build_regions failed: {rc}")
    n, nvo = int(cb.n_regions), int(cb.variants.n_variants)
    return RegionBatch(k, b.region_id[:n], b.contig[:n], b.start[:n], b.end[:n], b.var_off[:n * k + 1], b.position[:nvo],
                       b.variant_type[:nvo], b.zygosity[:nvo], b.raw_allele_space[:nvo], b.allele_off[:nvo], b.a0_len[:nvo],
                       b.a1_len[:nvo], b.allele_pool[:max(int(cb.variants.allele_pool_len), 1)])


def build_regions_bed(callsets, contig_lens, flank: int, bed=None, first_region_id: int = 0) -> RegionBatch:
    """RegionIterator restatement over several contigs and BED intervals (region_generation.rs:276-479) -> RegionBatch."""
    nv, k = callsets.n_variants, callsets.n_inputs
    b = RegionBatch(k, np.zeros(nv, np.uint64), np.zeros(nv, np.uint32), np.zeros(nv, np.uint32), np.zeros(nv, np.uint32),
                    np.zeros(nv * k + 1, np.uint64), np.zeros(nv, np.uint32), np.zeros(nv, np.uint8), np.zeros(nv, np.uint8),
                    np.zeros(nv, np.uint32), np.zeros(nv, np.uint32), np.zeros(nv, np.uint32), np.zeros(nv, np.uint32),
                    np.zeros(max(callsets.pool_len, 1), np.uint8))
    cs, cb = callsets.to_c(), b.to_c()
    vc = np.ascontiguousarray(callsets.variant_contig if callsets.variant_contig is not None and callsets.variant_contig.size else np.zeros(max(nv, 1), np.uint32),
                              dtype=np.uint32)
    lens = np.asarray(contig_lens, dtype=np.uint64)
    cbed = bed.to_c() if bed is not None else None
    fn = lib().orc_build_regions_bed
    fn.argtypes = [C.POINTER(abi.CallSets), C.POINTER(C.c_uint32), C.POINTER(abi.BedIntervals), C.POINTER(C.c_uint64), C.c_uint32, C.c_uint32,
                   C.c_uint64, C.POINTER(abi.RegionBatch)]
    rc = fn(C.byref(cs), abi.ptr(vc), C.byref(cbed) if cbed is not None else None, abi.ptr(lens), len(lens), flank, first_region_id, C.byref(cb))
    if rc != 0:
        raise RuntimeError(f"orc_build_regions_bed failed: {rc}")
    n, nvo = int(cb.n_regions), int(cb.variants.n_variants)
    return RegionBatch(k, b.region_id[:n], b.contig[:n], b.start[:n], b.end[:n], b.var_off[:n * k + 1], b.position[:nvo],
                       b.variant_type[:nvo], b.zygosity[:nvo], b.raw_allele_space[:nvo], b.allele_off[:nvo], b.a0_len[:nvo],
                       b.a1_len[:nvo], b.allele_pool[:max(int(cb.variants.allele_pool_len), 1)])


def vcf_parse(text: bytes, contig_names, sample_index=0, enable_trimming=True):
    """parse_variant / parse_genotype / get_variant_type restatement over VCF record lines -> (VcfTable, 0) or (None, (line, code))."""
    from aardvark_b200.ingest import VcfTable
    tab = VcfTable(2 * (text.count(b"\n") + 1), len(text))
    c = tab.to_c()
    names = (C.c_char_p * len(contig_names))(*[n.encode() for n in contig_names])
    code = C.c_int32(0)
    fn = lib().orc_vcf_parse
    fn.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32, C.c_int, C.POINTER(abi.VcfOut), C.POINTER(C.c_int32)]
    rc = fn(text, len(text), names, len(contig_names), sample_index, 1 if enable_trimming else 0, C.byref(c), C.byref(code))
    if rc != 0:
        return None, (rc - 1, int(code.value))
    return tab.finish(c), 0
