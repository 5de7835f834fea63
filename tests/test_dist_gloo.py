"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous region bins + the single
result gather.  The per-bin solve is done by the CPU oracle here (this is a test); on the GPU box
the same partition/gather code wraps the CUDA solver (bench.py, N > 1)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

import oracle_py as orc
from aardvark_b200 import abi, synth
from aardvark_b200.dist import gather_compare_outputs, partition_regions, region_cost_proxy


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref, batch = synth.workload_chr20(scale=0.01, seed=20)
    bins = partition_regions(batch, world)
    lo, hi = bins[rank]
    sub = batch.slice_regions(lo, hi)
    out = orc.compare_batch(sub, [ref], abi.CompareCfg(50, 0, 0, 0), n_threads=1)
    merged = gather_compare_outputs(out, sub.n_regions, sub.n_variants)
    if rank == 0:
        full = orc.compare_batch(batch, [ref], abi.CompareCfg(50, 0, 0, 0), n_threads=1)
        ok = (merged["n_regions"] == batch.n_regions and merged["n_variants"] == batch.n_variants
              and np.array_equal(merged["totals"], full.totals)
              and merged["totals_mask"] == int(full.totals_mask[0])
              and merged["solved"] == int(full.solved_blocks[0])
              and all(np.array_equal(merged[f], getattr(full, f)[:len(merged[f])])
                      for f in ("status", "ed1", "ed2", "type_mask", "var_expected", "var_observed", "var_class")))
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_contiguous_and_balanced():
    ref, batch = synth.workload_chr20(scale=0.02, seed=20)
    for world in (1, 2, 3, 8):
        bins = partition_regions(batch, world)
        assert bins[0][0] == 0 and bins[-1][1] == batch.n_regions
        assert all(bins[i][1] == bins[i + 1][0] for i in range(world - 1))
        cost = region_cost_proxy(batch)
        per = [cost[lo:hi].sum() for lo, hi in bins]
        assert max(per) <= 1.25 * (sum(per) / world) + cost.max()


def test_slice_regions_roundtrip():
    ref, batch = synth.workload_chr20(scale=0.01, seed=21)
    cfg = abi.CompareCfg(50, 0, 0, 0)
    full = orc.compare_batch(batch, [ref], cfg)
    lo, hi = batch.n_regions // 3, 2 * batch.n_regions // 3
    sub = batch.slice_regions(lo, hi)
    part = orc.compare_batch(sub, [ref], cfg)
    assert np.array_equal(part.status[:hi - lo], full.status[lo:hi])
    assert np.array_equal(part.region_metrics[:hi - lo], full.region_metrics[lo:hi])
    v0, v1 = int(batch.var_off[2 * lo]), int(batch.var_off[2 * hi])
    assert np.array_equal(part.var_class[:v1 - v0], full.var_class[v0:v1])


def test_two_rank_gather_matches_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
