"""Multi-GPU sharding of the compare/merge path.

Clusters are independent (docs/methods.md:17, src/main.rs:251), so the region_id-ordered
cluster list is cut into contiguous bins, one per rank, balanced by a cost proxy.  There is
no collective on the data path; afterwards ONE gather brings the per-rank summary counters
and per-variant annotation arrays to rank 0 (NCCL over NVLink; gloo in the CPU tests).
Result order is restored simply by concatenating the bins (src/main.rs:271 sorts by
region_id; bins are contiguous in region_id).
"""
from typing import List, Tuple

import numpy as np

from . import abi
from .batch import CompareOutputs, RegionBatch


def region_cost_proxy(batch: RegionBatch) -> np.ndarray:
    """Per-region work estimate: (N + 1) * window + sum over alleles of len^2 (SV tail)."""
    k = batch.n_inputs
    vo = batch.var_off.astype(np.int64)
    n_var = vo[k::k] - vo[0:-1:k]
    win = batch.end.astype(np.int64) - batch.start.astype(np.int64)
    al = np.maximum(batch.a0_len, batch.a1_len).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(al * al)])
    sq = csum[vo[k::k]] - csum[vo[0:-1:k]]
    return (n_var + 1) * win + sq


def partition_regions(batch: RegionBatch, world: int) -> List[Tuple[int, int]]:
    """Contiguous bins [lo, hi) per rank with approximately equal cost."""
    n = batch.n_regions
    if n == 0:
        return [(0, 0)] * world
    cost = np.cumsum(region_cost_proxy(batch).astype(np.float64))
    total = cost[-1]
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cost, total * r / world, side="left")))
    cuts.append(n)
    cuts = [min(max(c, 0), n) for c in cuts]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


_HEADER = 4  # int64 words: n_regions, n_variants, solved, errors


def _pack(out: CompareOutputs, n_regions: int, n_variants: int) -> np.ndarray:
    parts = [
        np.array([n_regions, n_variants, int(out.solved_blocks[0]), int(out.error_blocks[0])], dtype=np.int64).view(np.uint8),
        np.ascontiguousarray(out.totals).view(np.uint8).reshape(-1),
        out.totals_mask.view(np.uint8),
        out.status[:n_regions].view(np.uint8), out.ed1[:n_regions].view(np.uint8), out.ed2[:n_regions].view(np.uint8),
        out.type_mask[:n_regions].view(np.uint8),
        out.var_expected[:n_variants], out.var_observed[:n_variants], out.var_class[:n_variants],
    ]
    return np.concatenate(parts)


def _unpack(buf: np.ndarray):
    hdr = buf[:8 * _HEADER].view(np.int64)
    n, nv = int(hdr[0]), int(hdr[1])
    o = 8 * _HEADER
    res = {"n_regions": n, "n_variants": nv, "solved": int(hdr[2]), "errors": int(hdr[3])}
    tb = 8 * abi.N_GROUPS * abi.N_METRICS
    res["totals"] = buf[o:o + tb].view(np.uint64).reshape(abi.N_GROUPS, abi.N_METRICS).copy(); o += tb
    res["totals_mask"] = int(buf[o:o + 2].view(np.uint16)[0]); o += 2
    res["status"] = buf[o:o + 4 * n].view(np.int32); o += 4 * n
    res["ed1"] = buf[o:o + 4 * n].view(np.uint32); o += 4 * n
    res["ed2"] = buf[o:o + 4 * n].view(np.uint32); o += 4 * n
    res["type_mask"] = buf[o:o + 2 * n].view(np.uint16); o += 2 * n
    for f in ("var_expected", "var_observed", "var_class"):
        res[f] = buf[o:o + nv]; o += nv
    return res


_STAGE = {}   # (device, nbytes) -> (pinned host staging tensor, device tensor): reused across calls


def _staging(dev, nbytes, pinned):
    import torch
    key = (str(dev), int(nbytes))
    if key not in _STAGE:
        if len(_STAGE) > 16:
            _STAGE.clear()
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=pinned)
        _STAGE[key] = (host, torch.empty(nbytes, dtype=torch.uint8, device=dev) if dev.type == "cuda" else host)
    return _STAGE[key]


def gather_compare_outputs(out: CompareOutputs, n_regions: int, n_variants: int, dst: int = 0):
    """The single result gather of the multi-GPU path.  Every rank contributes its summary counters
    and per-region / per-variant arrays; rank `dst` returns the merged result (concatenated in rank ==
    region_id order, totals summed with wrapping u64 like the reference's AddAssign), others None.
    Payloads travel through pinned staging buffers that are kept between calls."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    rank = dist.get_rank()
    cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
    payload = _pack(out, n_regions, n_variants)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, torch.tensor([payload.size], dtype=torch.int64, device=dev))
    sizes = sizes.tolist()
    max_size = (max(sizes) + 15) & ~15
    host, send = _staging(dev, max_size, cuda)
    host.numpy()[:payload.size] = payload
    if cuda:
        send.copy_(host, non_blocking=True)
    if rank == dst:
        rhost, recv = _staging(dev, max_size * world, cuda)
        dist.gather(send, list(recv.view(world, max_size).unbind(0)), dst=dst)
        if cuda:
            rhost.copy_(recv, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        rbuf = rhost.numpy().reshape(world, max_size)
    else:
        dist.gather(send, None, dst=dst)
        return None
    parts = [_unpack(rbuf[r, :sizes[r]]) for r in range(world)]
    merged = {
        "n_regions": sum(p["n_regions"] for p in parts), "n_variants": sum(p["n_variants"] for p in parts),
        "solved": sum(p["solved"] for p in parts), "errors": sum(p["errors"] for p in parts),
        "totals_mask": int(np.bitwise_or.reduce([p["totals_mask"] for p in parts])),
    }
    tot = np.zeros((abi.N_GROUPS, abi.N_METRICS), dtype=np.uint64)
    for p in parts:
        tot += p["totals"]
    merged["totals"] = tot
    for f in ("status", "ed1", "ed2", "type_mask", "var_expected", "var_observed", "var_class"):
        merged[f] = np.concatenate([p[f] for p in parts])
    return merged
