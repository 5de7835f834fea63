"""Rebuild the reference's result structs from the flat output arrays (host side of the ABI)."""
from typing import List, Union

from . import abi
from .batch import CompareOutputs, MergeOutputs, RegionBatch
from .types import (Classification, CompareBenchmark, GroupTypeMetrics, MergeBenchmark, MergeClassification,
                    RegionError, SequenceBundle, VariantMetrics, VariantSource)


def unpack_compare(batch: RegionBatch, out: CompareOutputs) -> List[Union[CompareBenchmark, RegionError]]:
    """One CompareBenchmark (src/data_types/compare_benchmark.rs:9-33) per region, or the
    RegionError standing for the reference's per-region anyhow::Error (src/main.rs:259-262)."""
    res = []
    for r in range(batch.n_regions):
        rid = int(batch.region_id[r])
        st = int(out.status[r])
        if st != abi.ST_OK:
            res.append(RegionError(rid, st))
            continue
        t0, q0, q1 = (int(batch.var_off[2 * r + k]) for k in range(3))
        truth = [VariantMetrics(VariantSource.Truth, Classification(int(out.var_class[i])),
                                int(out.var_expected[i]), int(out.var_observed[i])) for i in range(t0, q0)]
        query = [VariantMetrics(VariantSource.Query, Classification(int(out.var_class[i])),
                                int(out.var_expected[i]), int(out.var_observed[i])) for i in range(q0, q1)]
        gm = GroupTypeMetrics(out.region_metrics[r], int(out.type_mask[r])) if out.region_metrics is not None else None
        bundle = None
        if out.seq_pool is not None and any(int(out.seq_len[r * 5 + s]) for s in range(5)):
            bundle = SequenceBundle(*[out.sequence(r, s).decode("latin-1") for s in range(5)])
        res.append(CompareBenchmark(rid, int(out.ed1[r]), int(out.ed2[r]), gm, truth, query, bundle))
    return res


def unpack_merge(batch: RegionBatch, out: MergeOutputs) -> List[Union[MergeBenchmark, RegionError]]:
    res = []
    for r in range(batch.n_regions):
        rid = int(batch.region_id[r])
        st = int(out.status[r])
        if st != abi.ST_OK:
            res.append(RegionError(rid, st))
            continue
        kind = int(out.classification[r])
        idx = tuple(int(x) for x in out.indices[r, :int(out.n_indices[r])])
        res.append(MergeBenchmark(rid, MergeClassification(kind, idx)))
    return res
