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
    """Contiguous bins [lo, hi) per rank with approximately equal cost: the library's own binning
    (avk_partition_regions, also used by avk_compare_batch_multi), so that one-process-per-GPU and
    one-process-many-GPUs runs cut the genome at the same places."""
    from .lib import partition_regions as _c_partition
    return _c_partition(batch, world)


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


_STAGE = {}   # (role, device, nbytes) -> (pinned host staging tensor, device tensor): reused across calls


def _staging(role, dev, nbytes, pinned):
    """Staging buffers are keyed by role (send / recv): with one rank both have the same size and must not alias."""
    import torch
    key = (role, str(dev), int(nbytes))
    if key not in _STAGE:
        if len(_STAGE) > 16:
            if dev.type == "cuda":
                torch.cuda.synchronize()          # a non-blocking copy may still read a buffer that is about to be dropped
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
    host, send = _staging("send", dev, max_size, cuda)
    host.numpy()[:payload.size] = payload
    if cuda:
        send.copy_(host, non_blocking=True)
    if rank == dst:
        rhost, recv = _staging("recv", dev, max_size * world, cuda)
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


# ---------------------------------------------------------------------------------------------------------------------
# Device-side gather (NCCL): the single collective of a one-process-per-GPU run.  Every rank's results are still in its
# GPU's memory (avk_compare_result_device); they travel GPU -> root GPU over NVLink as ONE grouped send/recv, land in
# the root's full-size device arrays at the bin offsets (bins are contiguous and known to every rank, so no size
# exchange is needed), and reach the host in one copy per array.
_FIELDS = (("status", 4, False), ("ed1", 4, False), ("ed2", 4, False), ("type_mask", 2, False),
           ("var_expected", 1, True), ("var_observed", 1, True), ("var_class", 1, True))
N_TOTALS = abi.N_GROUPS * abi.N_METRICS + 3


class _DevArray:
    """Wraps a raw device address so that torch can view it (CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _dev_tensor(ptr, nbytes, dev):
    import torch
    if nbytes == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device=dev)
    return torch.as_tensor(_DevArray(ptr, nbytes), device=dev)


class DeviceGather:
    """Preallocated buffers of the gather: the root's full-size device arrays and its pinned host mirrors."""

    def __init__(self, bins, var_bins, rank, dst=0):
        import torch
        self.bins, self.var_bins, self.rank, self.dst = list(bins), list(var_bins), rank, dst
        self.world = len(self.bins)
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.n = self.bins[-1][1]
        self.nv = self.var_bins[-1][1]
        self.full, self.host = {}, {}
        if rank == dst:
            for name, width, per_var in _FIELDS:
                nbytes = max((self.nv if per_var else self.n) * width, 1)
                self.full[name] = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
                self.host[name] = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            self.full["totals"] = torch.empty(self.world * N_TOTALS * 8, dtype=torch.uint8, device=self.dev)
            self.host["totals"] = torch.empty(self.world * N_TOTALS * 8, dtype=torch.uint8, pin_memory=True)

    def d2h_bytes(self):
        return sum(t.numel() for t in self.host.values())

    def gather(self, view: "abi.CompareDevView"):
        """view: this rank's avk_compare_dev_view.  Returns the merged result dict on the root, None elsewhere."""
        import torch
        import torch.distributed as dist
        lo, hi = self.bins[self.rank]
        v0, v1 = self.var_bins[self.rank]
        assert view.lo == lo and view.n_regions == hi - lo and view.v_base == v0 and view.n_variants == v1 - v0
        mine = {name: _dev_tensor(getattr(view, name), ((v1 - v0) if per_var else (hi - lo)) * width, self.dev)
                for name, width, per_var in _FIELDS}
        mine["totals"] = _dev_tensor(view.totals, N_TOTALS * 8, self.dev)
        ops = []
        if self.rank == self.dst:
            for r in range(self.world):
                rlo, rhi = self.bins[r]
                rv0, rv1 = self.var_bins[r]
                for name, width, per_var in _FIELDS:
                    a, b = ((rv0, rv1) if per_var else (rlo, rhi))
                    if b == a:
                        continue
                    dst_t = self.full[name][a * width:b * width]
                    if r == self.rank:
                        dst_t.copy_(mine[name])
                    else:
                        ops.append(dist.P2POp(dist.irecv, dst_t, r))
                tot_t = self.full["totals"][r * N_TOTALS * 8:(r + 1) * N_TOTALS * 8]
                if r == self.rank:
                    tot_t.copy_(mine["totals"])
                else:
                    ops.append(dist.P2POp(dist.irecv, tot_t, r))
        else:
            for name, width, per_var in _FIELDS:
                if mine[name].numel():
                    ops.append(dist.P2POp(dist.isend, mine[name], self.dst))
            ops.append(dist.P2POp(dist.isend, mine["totals"], self.dst))
        if ops:
            for req in dist.batch_isend_irecv(ops):      # ONE NCCL group: the gather
                req.wait()
        if self.rank != self.dst:
            torch.cuda.current_stream().synchronize()     # the library may overwrite its result arrays after this call
            return None
        for name in self.full:
            self.host[name].copy_(self.full[name], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        res = {"n_regions": self.n, "n_variants": self.nv}
        dt = {"status": np.int32, "ed1": np.uint32, "ed2": np.uint32, "type_mask": np.uint16}
        for name, width, per_var in _FIELDS:
            cnt = self.nv if per_var else self.n
            res[name] = self.host[name].numpy()[:cnt * width].view(dt.get(name, np.uint8))
        tot = self.host["totals"].numpy().view(np.uint64).reshape(self.world, N_TOTALS)
        nm = abi.N_GROUPS * abi.N_METRICS
        with np.errstate(over="ignore"):
            res["totals"] = tot[:, :nm].sum(axis=0, dtype=np.uint64).reshape(abi.N_GROUPS, abi.N_METRICS)   # wrapping u64
        res["totals_mask"] = int(np.bitwise_or.reduce(tot[:, nm]) & 0xffff)
        res["solved"] = int(tot[:, nm + 1].sum())
        res["errors"] = int(tot[:, nm + 2].sum())
        return res
