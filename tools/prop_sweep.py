#!/usr/bin/env python
"""Offline property sweep (CPU): random clusters through the host builds of the two scalar device solvers against the oracle,
like tests/test_properties.py but wider (more records, longer alleles) and as many examples as asked for.
    MAXV=8 MAXINS=10 [LONG=1] python tools/prop_sweep.py N_EXAMPLES SEED"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "oracle", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import hypothesis
from hypothesis import HealthCheck, given, settings, strategies as st, seed
from aardvark_b200 import abi
from aardvark_b200.types import PhasedZygosity, Variant, VariantType
import test_properties as P
import test_spec_search_host as SP
import test_thread_solver_host as TS

ZYGS = P.ZYGS
@st.composite
def big_clusters(draw):
    L = draw(st.integers(60, 220))
    alphabet = draw(st.sampled_from([b"ACGT", b"AC", b"A", b"ACG"]))
    ref = bytes(draw(st.lists(st.sampled_from(list(alphabet)), min_size=L, max_size=L)))
    sides = []
    for _ in range(2):
        n = draw(st.integers(0, int(os.environ.get("MAXV", "8"))))
        pos = sorted(draw(st.lists(st.integers(10, L - 30), min_size=n, max_size=n)))
        lst = []
        for p in pos:
            l0 = draw(st.sampled_from([1, 1, 1, 2, 3, 5, 9] + ([17, 26] if os.environ.get("LONG") else [])))
            a0 = ref[p:p + l0]
            kind = draw(st.integers(0, 4))
            if kind == 0: a1 = bytes([draw(st.sampled_from(list(b"ACGT")))]) + a0[1:]
            elif kind == 1: a1 = a0[:1] + bytes(draw(st.lists(st.sampled_from(list(alphabet)), min_size=1, max_size=int(os.environ.get("MAXINS", "10")))))
            elif kind == 2: a1 = a0[:1]
            elif kind == 3: a1 = bytes(draw(st.lists(st.sampled_from(list(b"ACGT")), min_size=1, max_size=3)))
            else: a1 = a0
            vt = (VariantType.Snv if len(a0) == 1 and len(a1) == 1 else VariantType.Insertion if len(a0) == 1 else VariantType.Deletion if len(a1) == 1 else VariantType.Indel)
            lst.append((Variant(0, vt, p, a0, a1, max(len(a0), len(a1))), draw(st.sampled_from(ZYGS))))
        sides.append(lst)
    if not sides[0] and not sides[1]:
        sides[0].append((Variant(0, VariantType.Snv, 30, ref[30:31], b"T" if ref[30:31] != b"T" else b"G", 1), PhasedZygosity.HomozygousAlternate))
    return ref, sides

N = int(sys.argv[1]); SEED = int(sys.argv[2])
cnt = {"n": 0, "ts": 0, "sp": 0}
@seed(SEED)
@settings(max_examples=N, deadline=None, suppress_health_check=list(HealthCheck), database=None)
@given(big_clusters(), st.sampled_from([50, 3, 1, 2]), st.sampled_from([0]))
def run(cl, mbf, shortcut):
    ref, sides = cl
    batch = P._batch(ref, sides)
    cfg = abi.CompareCfg(mbf, shortcut, 0, 0)
    _, rej, _ = TS.run_ts(batch, [ref], cfg)
    TS.check(batch, [ref], cfg)
    _, n_ok = SP.check_solve(batch, [ref], cfg)
    cnt["n"] += 1; cnt["ts"] += int(not rej[0]); cnt["sp"] += n_ok
run()
print("ok", cnt)
