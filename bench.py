#!/usr/bin/env python
"""bench.py -- variant clusters/sec through the compare solve phase (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path (alt_ed + per-cluster solve + summary reduction) over one
batch: the synthetic chr20-scale compare of BASELINE.json configs[1] (~150k variants per side,
~118k clusters).  Weak scaling: every rank solves its own chr20-sized contig (seed 20 + rank),
regions binned by contig across GPUs with no data-path collective; only the per-GPU counters and
per-variant annotation arrays are gathered (one NCCL gather) in the e2e leg.

  value  clusters/s with the batch already resident in HBM (device time, CUDA events on the
         library's stream, L2 flushed between steps)
  e2e    clusters/s through the C ABI call with HOST (pinned) buffers: H2D of the batch, kernels,
         D2H of status/ed/per-variant labels/summary counters inside the timed region
  --impl reference: the CPU restatement of the reference path (oracle "port", OpenMP over clusters,
         all host threads) on the same batch -- the reference is Rust and cannot be built here.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "variant_clusters_per_sec"
UNIT = "clusters/s"


def workload(rank, scale):
    from aardvark_b200 import synth
    return synth.workload_chr20(scale=scale, seed=20 + rank)


def algorithmic_bytes(batch):
    """Compulsory HBM bytes of one solve pass as laid out (DESIGN.md 'algorithmic bytes'): every input
    array read once, the reference window of every cluster read once, every output written once."""
    from aardvark_b200 import abi
    win = int((batch.end.astype(np.int64) - batch.start.astype(np.int64)).sum())
    out_bytes = batch.n_regions * (4 + 4 + 4 + 2 + 8 * abi.N_GROUPS * abi.N_METRICS) + 3 * batch.n_variants
    return batch.nbytes() + 4 * batch.n_variants + win + out_bytes


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="avk_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned_copy(arr):
    import torch
    t = torch.from_numpy(np.array(arr, copy=True)).pin_memory()
    return t.numpy()


def pin_batch(batch):
    from aardvark_b200.batch import RegionBatch
    f = lambda a: pinned_copy(a)
    return RegionBatch(batch.n_inputs, f(batch.region_id), f(batch.contig), f(batch.start), f(batch.end), f(batch.var_off),
                       f(batch.position), f(batch.variant_type), f(batch.zygosity), f(batch.raw_allele_space),
                       f(batch.allele_off), f(batch.a0_len), f(batch.a1_len), f(batch.allele_pool))


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port; the Rust crate cannot be built here)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as orc
    from aardvark_b200 import abi
    ref, batch = workload(0, args.scale)
    cfg = abi.CompareCfg(50, 0, 0, 0)
    threads = orc.num_threads()
    for _ in range(args.warmup):
        orc.compare_batch(batch, [ref], cfg, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.compare_batch(batch, [ref], cfg, n_threads=threads)
    dt = time.perf_counter() - t0
    v = batch.n_regions * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"chr20-scale synthetic compare (BASELINE configs[1]), scale={args.scale}, seed=20",
                   "regions_per_step": batch.n_regions, "variants_per_step": batch.n_variants, "max_branch_factor": 50},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"whole batch ({batch.n_regions} clusters) x {args.steps} steps, OpenMP dynamic over clusters"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the chr20-scale workload (tests only; default = full)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from aardvark_b200 import abi
    from aardvark_b200.batch import CompareOutputs
    from aardvark_b200.dist import gather_compare_outputs
    from aardvark_b200.lib import Solver
    from aardvark_b200.types import CompareConfig

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: aardvark_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries ONE JSON line: NCCL's version banner (NCCL_DEBUG=VERSION) would land there too
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ref, batch = workload(rank, args.scale)
    solver = Solver(local_rank)
    solver.set_reference([ref])
    cfg = CompareConfig(enable_sequences=False)
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident leg: `value` -------------------------------------------------------
    solver.upload(batch)
    for _ in range(args.warmup):
        solver.run_resident(cfg)
    launches0 = solver.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    dev_ms, search_ms = 0.0, 0.0
    for _ in range(args.steps):
        l2_flush.zero_()
        torch.cuda.synchronize()
        solver.run_resident(cfg)           # returns after its CUDA events (own stream) have completed
        tm = solver.last_timings_ms()
        dev_ms += tm["total"]
        search_ms += tm["search"]
    barrier()
    launches = solver.launch_count() - launches0
    work = solver.last_work()
    resident_out = solver.download(CompareOutputs(batch, region_metrics=False))

    # ---------------- e2e leg: host buffers through the C ABI ----------------------------------------
    pbatch = pin_batch(batch)
    out = CompareOutputs(pbatch, region_metrics=False)
    for f in ("status", "ed1", "ed2", "type_mask", "var_expected", "var_observed", "var_class"):
        setattr(out, f, pinned_copy(getattr(out, f)))
    for _ in range(3):
        solver.compare_batch(pbatch, cfg, out=out)
        if world > 1:
            gather_compare_outputs(out, pbatch.n_regions, pbatch.n_variants)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        solver.compare_batch(pbatch, cfg, out=out)
        if world > 1:
            gather_compare_outputs(out, pbatch.n_regions, pbatch.n_variants)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None
    assert out.diff(resident_out) == [], "e2e and resident outputs differ"
    h2d = pbatch.nbytes()
    d2h = sum(getattr(out, f).nbytes for f in ("status", "ed1", "ed2", "type_mask", "var_expected", "var_observed", "var_class")) \
        + out.totals.nbytes + 24

    # ---------------- aggregate over ranks: max time, summed clusters --------------------------------
    stats = torch.tensor([dev_ms, e2e_s * 1e3, search_ms], dtype=torch.float64, device="cuda")
    counts = torch.tensor([batch.n_regions, batch.n_variants, int(out.solved_blocks[0]), int(out.error_blocks[0]), h2d, d2h],
                          dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max, search_ms_max = (float(x) for x in stats.tolist())
    n_regions, n_variants, solved, errors, h2d_all, d2h_all = (int(x) for x in counts.tolist())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        alg_bytes = algorithmic_bytes(batch)
        k_ms = search_ms / args.steps                       # rank 0's dominant kernel (k_compare, all tiers)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        int_peak = solver.int_peak_ops_per_s()
        int_ops = 6 * work["cells"] + 4 * ((work["matched_bases"] + 15) // 16)
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "k_compare_traffic.json")))
            if abs(prof.get("scale", -1) - args.scale) < 1e-9:
                traffic = prof.get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": n_regions * args.steps / (dev_ms_max * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"chr20-scale synthetic compare (BASELINE configs[1]) per GPU, scale={args.scale}, seed=20+rank",
                       "regions_per_step": n_regions, "variants_per_step": n_variants, "max_branch_factor": 50,
                       "l2": "flushed between timed steps (256 MiB write)", "solved_blocks": solved, "error_blocks": errors,
                       "partition": "one contig bin per GPU, no data-path collective; single gather of results in e2e"},
            "e2e": {"value": n_regions * args.steps / (e2e_ms_max * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all, "ms_per_step": e2e_ms_max / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "kernel": "k_compare", "kernel_ms": k_ms, "algorithmic_bytes": alg_bytes,
                         "peak_source": peak_src,
                         "note": "integer DP: the binding roof is the INT32 ALU pipe / latency, see int_roofline"},
            # executed on the device (closed forms and pruned searches do less than the reference algorithm);
            # replaced below by the reference algorithm's own counts when the CPU leg runs
            "int_roofline": {"achieved_gops": int_ops / (k_ms * 1e-3) / 1e9, "peak_gops": int_peak / 1e9,
                             "frac": int_ops / (k_ms * 1e-3) / int_peak, "algorithmic_int_ops": int_ops,
                             "ops_counted_by": "device counters (work actually executed)",
                             "cells": work["cells"], "matched_bases": work["matched_bases"],
                             "search_pops": work["search_pops"], "exact_pops": work["exact_pops"],
                             "peak_source": "measured live (avk_int_peak: add/max/xor chains)"},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle_py as orc
            from aardvark_b200.lib import compare_cfg
            threads = orc.num_threads()
            best = None
            cpu_out = None
            reps = 3
            for _ in range(reps):
                t0 = time.perf_counter()
                cpu_out = orc.compare_batch(batch, [ref], compare_cfg(cfg), n_threads=threads, region_metrics=False)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            # algorithmic work = what the reference algorithm computes, counted by the CPU restatement
            _, cw = orc.compare_batch(batch, [ref], compare_cfg(cfg), n_threads=threads, region_metrics=False, work=True)
            ref_ops = 6 * cw["cells"] + 4 * ((cw["matched_bases"] + 15) // 16)
            ir = line["int_roofline"]
            ir.update({"device_executed": {k: ir[k] for k in ("algorithmic_int_ops", "cells", "matched_bases", "search_pops", "exact_pops")},
                       "achieved_gops": ref_ops / (k_ms * 1e-3) / 1e9, "frac": ref_ops / (k_ms * 1e-3) / int_peak,
                       "algorithmic_int_ops": ref_ops, "cells": cw["cells"], "matched_bases": cw["matched_bases"],
                       "search_pops": cw["search_pops"], "exact_pops": cw["exact_pops"],
                       "ops_counted_by": "CPU restatement of the reference algorithm (6*cells + 4*ceil(matched/16))"})
            line["cpu_baseline"] = {"value": batch.n_regions / best, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"rank 0's whole batch ({batch.n_regions} clusters), best of {reps}, "
                                              "OpenMP dynamic over clusters, solve phase only",
                                    "matches_gpu_bit_exact": out.diff(cpu_out) == []}
        print(json.dumps(line))
    solver.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
