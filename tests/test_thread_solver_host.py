"""Host build of the thread-per-cluster solver (aardvark_b200/csrc/avk_thread_solver.cuh, compiled for the CPU by
tests/ts_host.cpp) against the CPU oracle: every cluster the fast path accepts must agree bit for bit; what it rejects
is solved by the warp kernels on the GPU (checked by the GPU parity tests)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_py as orc
from aardvark_b200 import abi, synth
from aardvark_b200.batch import CompareOutputs

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libts_host.so")
SRCS = [os.path.join(HERE, "ts_host.cpp"), os.path.join(HERE, "..", "aardvark_b200", "csrc", "avk_thread_solver.cuh"),
        os.path.join(HERE, "..", "aardvark_b200", "csrc", "avk_layout.h")]


def ts_lib():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(s) for s in SRCS):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", SO, SRCS[0]])
    lib = C.CDLL(SO)
    lib.ts_compare_batch.argtypes = [C.POINTER(abi.RegionBatch), C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_uint64), C.c_uint32,
                                     C.POINTER(abi.CompareCfg), C.POINTER(abi.CompareOut), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)]
    return lib


def run_ts(batch, contigs, cfg):
    lib = ts_lib()
    out = CompareOutputs(batch)
    arrs, ptrs, lens = orc._contig_args(contigs)
    rej = np.ones(max(batch.n_regions, 1), dtype=np.uint8)
    stats = np.zeros(8, dtype=np.uint64)
    cb, co = batch.to_c(), out.to_c()
    rc = lib.ts_compare_batch(C.byref(cb), ptrs, lens, len(arrs), C.byref(cfg), C.byref(co), abi.ptr(rej), abi.ptr(stats))
    assert rc == 0
    return out, rej[:batch.n_regions].astype(bool), stats


def check(batch, contigs, cfg=None, min_accept=0.0):
    cfg = cfg or abi.CompareCfg(50, 0, 0, 0)
    ts, rej, stats = run_ts(batch, contigs, cfg)
    cpu = orc.compare_batch(batch, contigs, cfg, n_threads=orc.num_threads())
    ok = ~rej
    n = batch.n_regions
    for f in ("status", "ed1", "ed2", "type_mask"):
        a, b = getattr(ts, f)[:n][ok], getattr(cpu, f)[:n][ok]
        bad = np.nonzero(a != b)[0]
        assert bad.size == 0, (f, np.nonzero(ok)[0][bad[:5]], a[bad[:5]], b[bad[:5]])
    bad = np.nonzero((ts.region_metrics[:n][ok] != cpu.region_metrics[:n][ok]).any(axis=(1, 2)))[0]
    assert bad.size == 0, ("region_metrics", np.nonzero(ok)[0][bad[:5]])
    k = batch.n_inputs
    vo = batch.var_off.astype(np.int64)
    vmask = np.repeat(ok, (vo[k::k] - vo[0:-1:k]))
    for f in ("var_expected", "var_observed", "var_class"):
        a, b = getattr(ts, f)[:batch.n_variants][vmask], getattr(cpu, f)[:batch.n_variants][vmask]
        assert np.array_equal(a, b), f
    frac = float(ok.mean()) if n else 1.0
    assert frac >= min_accept, f"fast path accepted only {frac:.3f} of the clusters"
    return frac, stats


@pytest.mark.parametrize("seed", [20, 21, 22, 23])
def test_thread_solver_chr20_shape(seed):
    ref, batch = synth.workload_chr20(scale=0.05, seed=seed)
    frac, stats = check(batch, [ref], min_accept=0.97)


def test_thread_solver_branch_factors():
    ref, batch = synth.workload_chr20(scale=0.01, seed=5)
    for mbf in (50, 2, 1, 7):
        check(batch, [ref], abi.CompareCfg(mbf, 0, 0, 0))


def test_thread_solver_dense_clusters():
    p = synth.SynthParams(n_variants=3000, dense_frac=0.9, dense_mean=6.0, het_frac=0.95, phased_frac=0.2,
                          p_repr=0.05, p_gt_err=0.05, p_fn=0.05, p_fp=0.05)
    ref, batch = synth.workload_compare(60_000, p, seed=11)
    check(batch, [ref])


def test_thread_solver_indel_heavy():
    p = synth.SynthParams(n_variants=4000, snv_frac=0.3, indel_geom_mean=6.0, p_repr=0.08, p_gt_err=0.05, p_fn=0.05, p_fp=0.05)
    ref, batch = synth.workload_compare(600_000, p, seed=13)
    check(batch, [ref])


def test_thread_solver_sv_shape_mostly_rejected_but_exact():
    p = synth.SynthParams(n_variants=300, sv_events=40, sv_max=800, flank=1000)
    ref, batch = synth.workload_compare(400_000, p, seed=4)
    check(batch, [ref])


def test_thread_solver_exact_gt_expansion_cap():
    """AVK_ST_TIMEOUT: with a tiny cap the searches that the (pruned) fast path runs and the oracle's agree on which regions
    time out whenever the first search of a region trips it; here every region's status must match."""
    p = synth.SynthParams(n_variants=1500, dense_frac=0.9, dense_mean=6.0, het_frac=0.95, phased_frac=0.2,
                          p_repr=0.05, p_gt_err=0.08, p_fn=0.08, p_fp=0.08)
    ref, batch = synth.workload_compare(30_000, p, seed=17)
    cfg = abi.CompareCfg(50, 0, 0, 0, 3)
    ts, rej, stats = run_ts(batch, [ref], cfg)
    cpu = orc.compare_batch(batch, [ref], cfg, n_threads=orc.num_threads())
    n = batch.n_regions
    ok = ~rej
    timed_out = cpu.status[:n] == abi.ST_TIMEOUT
    assert timed_out.sum() > 0
    # the fast path may prune a search the oracle runs to the cap (documented); where it also times out, outputs agree
    both = ok & (ts.status[:n] == abi.ST_TIMEOUT)
    assert both.sum() > 0 and (timed_out[both]).all()
    same = ok & (ts.status[:n] == cpu.status[:n])
    assert (ts.region_metrics[:n][same] == cpu.region_metrics[:n][same]).all()
