"""Microbenchmark of the batched global edit distance (avk_wfa_ed_batch -> k_wfa_ed_cta): long pairs with edit distances of
5-10 k, i.e. the wide-wavefront regime of BASELINE configs[3].  Prints one JSON line: wavefront cells per second (cells =
sum of (ED + 1)^2), per SM, and the fraction of the measured INT32 peak at 6 integer ops per cell + 4 per 16
matched bases."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aardvark_b200 import synth
from aardvark_b200.lib import Solver

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 296
rng = np.random.default_rng(5)
pairs = []
for i in range(n_pairs):
    a = synth.ACGT[rng.integers(0, 4, size=14000)]
    kind = i % 3
    if kind == 0:      # unrelated sequence of 10 kbp: ED ~ 5.5 k
        a = a[:10000]; b = synth.ACGT[rng.integers(0, 4, size=10000)]
    elif kind == 1:    # a 7 kbp deletion inside 14 kbp
        b = np.concatenate([a[:3000], a[10000:]])
    else:              # a 9 kbp insertion into 5 kbp, plus 1 % substitutions
        b = a.copy(); a = np.concatenate([a[:2000], a[11000:]])
        idx = rng.integers(0, b.size, size=140); b[idx] = synth.ACGT[rng.integers(0, 4, size=140)]
    pairs.append((a.tobytes(), b.tobytes()))
s = Solver(0)
for _ in range(2):
    t0 = time.perf_counter(); ed = s.wfa_ed_batch(pairs); wall = time.perf_counter() - t0
tm = s.last_timings_ms(); w = s.last_work()
cells = int(((ed.astype(np.int64) + 1) ** 2).sum())
k_ms = tm["search"]
peak = s.int_peak_ops_per_s()
ops = 6 * cells + 4 * ((w["matched_bases"] + 15) // 16)
print(json.dumps({"metric": "wavefront_cells_per_sec", "kernel": "k_wfa_ed_cta", "pairs": n_pairs, "ed_min": int(ed.min()), "ed_max": int(ed.max()),
                  "cells": cells, "kernel_ms": k_ms, "wall_ms_with_copies": wall * 1e3, "gcells_per_s": cells / k_ms / 1e6,
                  "gcells_per_s_per_sm": cells / k_ms / 1e6 / 148, "int_roofline": {"ops": ops, "peak_gops": peak / 1e9, "frac": ops / (k_ms * 1e-3) / peak},
                  "device_counted_cells": w["cells"]}))
