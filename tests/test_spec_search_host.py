"""Host build of the speculative dense-cluster search (aardvark_b200/csrc/avk_spec_search.cuh, compiled for the CPU by
tests/sp_host.cpp, 32 simulated lanes) against the CPU oracle's optimize_sequences: every cluster the batched search accepts
must return the same equal-best results -- same order, allele assignments, edit distances, skipped-variant costs.  What it
rejects is solved by the team stage from scratch on the GPU (checked by the GPU parity tests)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_py as orc
from aardvark_b200 import abi, synth

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libsp_host.so")
SRCS = [os.path.join(HERE, "sp_host.cpp"), os.path.join(HERE, "host_digest.h"),
        os.path.join(HERE, "..", "aardvark_b200", "csrc", "avk_spec_search.cuh"),
        os.path.join(HERE, "..", "aardvark_b200", "csrc", "avk_layout.h")]
RESCAP = 64
ZYG_OF = {(0, 0): abi.ZYG_HOM_REF, (0, 1): abi.ZYG_PHASED_HET01, (1, 0): abi.ZYG_PHASED_HET10, (1, 1): abi.ZYG_HOM_ALT}


def sp_lib():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(s) for s in SRCS):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", SO, SRCS[0]])
    lib = C.CDLL(SO)
    lib.sp_search_batch.argtypes = [C.POINTER(abi.RegionBatch), C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_uint64), C.c_uint32,
                                    C.c_uint32, C.c_uint32, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)]
    lib.sp_solve_batch.argtypes = [C.POINTER(abi.RegionBatch), C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_uint64), C.c_uint32,
                                   C.POINTER(abi.CompareCfg), C.c_uint32, C.POINTER(abi.CompareOut), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)]
    return lib


def run_sp(batch, contigs, mbf=50, min_n=1, score=False):
    lib = sp_lib()
    n = max(batch.n_regions, 1)
    status = np.zeros(n, dtype=np.uint8)
    nres = np.zeros(n, dtype=np.uint32)
    res = np.zeros((n, RESCAP, 8), dtype=np.uint32)
    tmask = np.zeros(n, dtype=np.uint32)
    stats = np.zeros(10, dtype=np.uint64)
    arrs, ptrs, lens = orc._contig_args(contigs)
    cb = batch.to_c()
    vexp = np.full(max(batch.n_variants, 1), 255, dtype=np.uint8)
    vobs = np.full(max(batch.n_variants, 1), 255, dtype=np.uint8)
    ed = np.zeros((n, 2), dtype=np.uint32)
    rc = lib.sp_search_batch(C.byref(cb), ptrs, lens, len(arrs), mbf, min_n, abi.ptr(status), abi.ptr(nres), abi.ptr(res), abi.ptr(tmask), abi.ptr(stats),
                             abi.ptr(vexp) if score else None, abi.ptr(vobs) if score else None, abi.ptr(ed) if score else None)
    assert rc == 0
    if score:
        return status, nres, res, tmask, stats, vexp, vobs, ed
    return status, nres, res, tmask, stats


def check(batch, contigs, mbf=50, min_n=1, min_accept=0.0):
    status, nres, res, tmask, stats = run_sp(batch, contigs, mbf, min_n)
    tried = np.nonzero(status != 2)[0]
    acc = np.nonzero(status == 0)[0]
    assert tried.size > 0
    assert acc.size >= min_accept * tried.size, (acc.size, tried.size)
    for r in acc.tolist():
        one = batch.slice_regions(r, r + 1)
        c = int(batch.contig[r])
        st, ref_res = orc.optimize_sequences(one, bytes(contigs[c]) if not isinstance(contigs[c], np.ndarray) else contigs[c].tobytes(), mbf, RESCAP)
        assert st == 0, (r, st)
        assert len(ref_res) == int(nres[r]), (r, len(ref_res), int(nres[r]))
        nt = int(one.var_off[1] - one.var_off[0])
        nq = int(one.var_off[2] - one.var_off[1])
        for i, rr in enumerate(ref_res):
            a1, a2, ed1, ed2, tv1, tv2, qv1, qv2 = (int(x) for x in res[r, i])
            assert (ed1, ed2, tv1, tv2, qv1, qv2) == (rr["ed1"], rr["ed2"], rr["truth_vs1"], rr["truth_vs2"], rr["query_vs1"], rr["query_vs2"]), (r, i)
            tz, qz = [], []
            for oi in range(nt + nq):
                z = ZYG_OF[((a1 >> oi) & 1, (a2 >> oi) & 1)]
                (tz if (int(tmask[r]) >> oi) & 1 else qz).append(z)
            assert tz == [int(x) for x in rr["truth_zygosity"]] and qz == [int(x) for x in rr["query_zygosity"]], (r, i)
    return stats, acc.size, tried.size


def test_spec_search_chr20_shaped_all_clusters():
    ref, batch = synth.workload_chr20(scale=0.01, seed=5)
    stats, acc, tried = check(batch, [ref], min_accept=0.95)
    assert stats[6] + 32 * stats[7] <= 28 * 1024      # per-warp shared memory: 8 warps per SM


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_spec_search_dense_clusters(seed):
    p = synth.SynthParams(n_variants=900, dense_frac=0.8, dense_mean=6.0, het_frac=0.9, phased_frac=0.3, p_repr=0.05, p_gt_err=0.05,
                          p_fn=0.05, p_fp=0.05)
    ref, batch = synth.workload_compare(30_000, p, seed=seed)
    stats, acc, tried = check(batch, [ref], min_n=6, min_accept=0.5)
    assert stats[5] > 1            # batches with more than one lane did occur


@pytest.mark.parametrize("mbf", [1, 2, 3, 7])
def test_spec_search_small_branch_quota(mbf):
    """max_branch_factor small enough that the per-depth quota drops nodes: batches are cut to the quota-safe prefix."""
    p = synth.SynthParams(n_variants=500, dense_frac=0.7, dense_mean=8.0, het_frac=0.9, phased_frac=0.2, p_gt_err=0.05, p_fn=0.05, p_fp=0.05)
    ref, batch = synth.workload_compare(20_000, p, seed=21)
    check(batch, [ref], mbf=mbf, min_n=3, min_accept=0.5)


def check_scoring(batch, contigs, mbf=50, min_n=1, min_accept=0.0):
    """Lane-parallel exact-GT scoring of the equal-best results against the oracle's full solve: chosen solution's edit
    distances and every variant's expected / observed ALT copies."""
    status, nres, res, tmask, stats, vexp, vobs, ed = run_sp(batch, contigs, mbf, min_n, score=True)
    cpu = orc.compare_batch(batch, contigs, abi.CompareCfg(mbf, 0, 0, 0), n_threads=orc.num_threads())
    acc = np.nonzero(status == 0)[0]
    tried = np.nonzero(status != 2)[0]
    assert acc.size >= min_accept * max(tried.size, 1), (acc.size, tried.size)
    vo = batch.var_off.astype(np.int64)
    for r in acc.tolist():
        if cpu.status[r] != 0:          # regions the full solve fails later (e.g. TP underflow) still searched and scored fine
            continue
        assert (int(ed[r, 0]), int(ed[r, 1])) == (int(cpu.ed1[r]), int(cpu.ed2[r])), r
        v0, v1 = int(vo[2 * r]), int(vo[2 * r + 2])
        assert np.array_equal(vexp[v0:v1], cpu.var_expected[v0:v1]) and np.array_equal(vobs[v0:v1], cpu.var_observed[v0:v1]), r
    return stats, acc.size, tried.size


def test_spec_scoring_chr20_shaped():
    ref, batch = synth.workload_chr20(scale=0.01, seed=6)
    check_scoring(batch, [ref], min_accept=0.95)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_spec_scoring_dense_clusters(seed):
    p = synth.SynthParams(n_variants=900, dense_frac=0.8, dense_mean=6.0, het_frac=0.9, phased_frac=0.3, p_repr=0.05, p_gt_err=0.05,
                          p_fn=0.05, p_fp=0.05)
    ref, batch = synth.workload_compare(30_000, p, seed=seed)
    check_scoring(batch, [ref], min_n=6, min_accept=0.5)


def check_solve(batch, contigs, cfg=None, min_n=1, min_accept=0.0):
    """Search + scoring + lane-parallel final metrics + commit: every output array against the oracle."""
    from aardvark_b200.batch import CompareOutputs
    cfg = cfg or abi.CompareCfg(50, 0, 0, 0)
    lib = sp_lib()
    out = CompareOutputs(batch)
    arrs, ptrs, lens = orc._contig_args(contigs)
    rej = np.ones(max(batch.n_regions, 1), dtype=np.uint8)
    stats = np.zeros(4, dtype=np.uint64)
    cb, co = batch.to_c(), out.to_c()
    assert lib.sp_solve_batch(C.byref(cb), ptrs, lens, len(arrs), C.byref(cfg), min_n, C.byref(co), abi.ptr(rej), abi.ptr(stats)) == 0
    cpu = orc.compare_batch(batch, contigs, cfg, n_threads=orc.num_threads())
    n = batch.n_regions
    ok = ~rej[:n].astype(bool)
    assert ok.sum() >= min_accept * max(int(stats[0]), 1), (int(ok.sum()), int(stats[0]))
    for f in ("status", "ed1", "ed2", "type_mask"):
        a, b = getattr(out, f)[:n][ok], getattr(cpu, f)[:n][ok]
        bad = np.nonzero(a != b)[0]
        assert bad.size == 0, (f, np.nonzero(ok)[0][bad[:5]], a[bad[:5]], b[bad[:5]])
    bad = np.nonzero((out.region_metrics[:n][ok] != cpu.region_metrics[:n][ok]).any(axis=(1, 2)))[0]
    assert bad.size == 0, ("region_metrics", np.nonzero(ok)[0][bad[:5]])
    vo = batch.var_off.astype(np.int64)
    vmask = np.repeat(ok, (vo[2::2] - vo[0:-2:2]))
    for f in ("var_expected", "var_observed", "var_class"):
        assert np.array_equal(getattr(out, f)[:batch.n_variants][vmask], getattr(cpu, f)[:batch.n_variants][vmask]), f
    return stats, int(ok.sum())


def test_spec_solve_chr20_shaped():
    ref, batch = synth.workload_chr20(scale=0.02, seed=8)
    check_solve(batch, [ref], min_accept=0.95)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_spec_solve_dense_clusters(seed):
    p = synth.SynthParams(n_variants=900, dense_frac=0.8, dense_mean=6.0, het_frac=0.9, phased_frac=0.3, p_repr=0.05, p_gt_err=0.05,
                          p_fn=0.05, p_fp=0.05)
    ref, batch = synth.workload_compare(30_000, p, seed=seed)
    check_solve(batch, [ref], min_n=4, min_accept=0.5)


def test_spec_solve_indel_heavy():
    """Mostly indels and MNP-like representation differences: the per-type filtered alignments of the metrics are exercised."""
    p = synth.SynthParams(n_variants=1200, snv_frac=0.3, dense_frac=0.6, dense_mean=10.0, het_frac=0.8, phased_frac=0.5, p_repr=0.2, p_gt_err=0.05,
                          p_fn=0.05, p_fp=0.05)
    ref, batch = synth.workload_compare(40_000, p, seed=3)
    stats, acc = check_solve(batch, [ref], min_n=2, min_accept=0.5)
    assert stats[2] > 0            # alignments without a closed form did occur
