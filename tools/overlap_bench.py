#!/usr/bin/env python
"""Experiment: the whole-genome compare through avk_compare_batch (pinned host buffers in, pinned results out) with the
batch cut into N concurrent lanes on ONE GPU (AVK_LANES), and with several whole genomes in flight from several host
threads (one context each).  Prints wall milliseconds per genome."""
import json, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

FIELDS = ("status", "ed1", "ed2", "type_mask", "var_expected", "var_observed", "var_class")


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    lane_list = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,2,3,4,6").split(",")]
    flight_list = [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "").split(",") if x]
    refs, batch, desc = bench.workload("wgs", scale, bench.host_cores())
    import torch
    from aardvark_b200.batch import CompareOutputs
    from aardvark_b200.lib import Solver
    from aardvark_b200.types import CompareConfig
    cfg = CompareConfig(enable_sequences=False)
    pbatch = bench.pin_batch(batch)

    def new_out():
        out = CompareOutputs(pbatch, region_metrics=False)
        for f in FIELDS:
            setattr(out, f, bench.pinned_copy(getattr(out, f)))
        return out

    res = {}
    ref_out = None
    os.environ["AVK_LANE_MIN_REGIONS"] = "1000"
    for n in lane_list:
        os.environ["AVK_LANES"] = str(n)
        s = Solver(0)
        s.set_reference(refs)
        out = new_out()
        for _ in range(3):
            s.compare_batch(pbatch, cfg, out=out)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            s.compare_batch(pbatch, cfg, out=out)
        torch.cuda.synchronize()
        res[f"one_genome_lanes_{n}"] = {"ms_per_genome": (time.perf_counter() - t0) * 1e3 / steps}
        if ref_out is None:
            ref_out = out
        else:
            assert out.diff(ref_out) == [], f"lanes {n} differ"
        s.close()
    for spec in flight_list:                     # e.g. 23 = 2 genomes in flight, 3 lanes each
        nf, nl = spec // 10, spec % 10
        os.environ["AVK_LANES"] = str(nl)
        lanes = []
        for _ in range(nf):
            s = Solver(0)
            s.set_reference(refs)
            lanes.append((s, new_out()))
        for s, out in lanes:
            for _ in range(2):
                s.compare_batch(pbatch, cfg, out=out)

        def work(i):
            s, out = lanes[i]
            for _ in range(steps):
                s.compare_batch(pbatch, cfg, out=out)
        th = [threading.Thread(target=work, args=(i,)) for i in range(nf)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize()
        res[f"in_flight_{nf}_lanes_{nl}"] = {"ms_per_genome": (time.perf_counter() - t0) * 1e3 / (nf * steps)}
        for s, out in lanes:
            assert out.diff(ref_out) == []
            s.close()
    print(json.dumps({"regions": int(batch.n_regions), "results": res}))


if __name__ == "__main__":
    main()
