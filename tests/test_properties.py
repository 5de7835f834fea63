"""Property tests (hypothesis) on the CPU: randomly drawn clusters -- overlapping records, repeated positions, every zygosity,
SNVs / insertions / deletions / indels on a low-complexity reference -- through the host builds of the two scalar device
solvers (thread per cluster: tests/ts_host.cpp; speculative, 32 simulated lanes: tests/sp_host.cpp) against the CPU oracle,
and the writers' f64 formatting against shortest-round-trip digits."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from aardvark_b200 import abi
from aardvark_b200.batch import RegionBatch
from aardvark_b200.types import Coordinates, CompareRegion, PhasedZygosity, Variant, VariantType
from aardvark_b200.writers import SummaryWriter
import test_spec_search_host as SP
import test_thread_solver_host as TS

SEEN = {"n": 0, "ts": 0, "sp": 0}
ZYGS = [PhasedZygosity.UnphasedHeterozygous, PhasedZygosity.PhasedHet01, PhasedZygosity.PhasedHet10, PhasedZygosity.HomozygousAlternate]


@st.composite
def clusters(draw):
    L = draw(st.integers(60, 160))
    alphabet = draw(st.sampled_from([b"ACGT", b"AC", b"A"]))          # low-complexity references make shifted representations equivalent
    ref = bytes(draw(st.lists(st.sampled_from(list(alphabet)), min_size=L, max_size=L)))
    sides = []
    for _ in range(2):
        n = draw(st.integers(0, 5))
        pos = sorted(draw(st.lists(st.integers(10, L - 20), min_size=n, max_size=n)))
        lst = []
        for p in pos:
            l0 = draw(st.sampled_from([1, 1, 1, 2, 3]))
            a0 = ref[p:p + l0]
            kind = draw(st.integers(0, 3))
            if kind == 0:
                a1 = bytes([draw(st.sampled_from(list(b"ACGT")))]) + a0[1:]                       # substitution (may equal the reference base)
            elif kind == 1:
                a1 = a0[:1] + bytes(draw(st.lists(st.sampled_from(list(alphabet)), min_size=1, max_size=4)))   # anchored insertion-like
            elif kind == 2:
                a1 = a0[:1]                                                                     # deletion (or a no-op SNV when l0 == 1)
            else:
                a1 = bytes(draw(st.lists(st.sampled_from(list(b"ACGT")), min_size=1, max_size=3)))   # arbitrary replacement
            vt = (VariantType.Snv if len(a0) == 1 and len(a1) == 1 else VariantType.Insertion if len(a0) == 1 else
                  VariantType.Deletion if len(a1) == 1 else VariantType.Indel)
            lst.append((Variant(0, vt, p, a0, a1, max(len(a0), len(a1))), draw(st.sampled_from(ZYGS))))
        sides.append(lst)
    if not sides[0] and not sides[1]:
        sides[0].append((Variant(0, VariantType.Snv, 30, ref[30:31], b"T" if ref[30:31] != b"T" else b"G", 1), PhasedZygosity.HomozygousAlternate))
    return ref, sides


def _batch(ref, sides):
    region = CompareRegion(0, Coordinates("c", 0, len(ref)), [v for v, _ in sides[0]], [z for _, z in sides[0]], [v for v, _ in sides[1]], [z for _, z in sides[1]])
    return RegionBatch.from_compare_regions([region], {"c": 0})


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(clusters(), st.sampled_from([50, 3, 1]))
def test_random_clusters_both_scalar_solvers_vs_oracle(cl, mbf):
    ref, sides = cl
    batch = _batch(ref, sides)
    cfg = abi.CompareCfg(mbf, 0, 0, 0)
    _, rej, _ = TS.run_ts(batch, [ref], cfg)
    TS.check(batch, [ref], cfg)                         # every cluster the thread solver accepts agrees bit for bit
    _, n_ok = SP.check_solve(batch, [ref], cfg)         # likewise the speculative solver (search, scoring, metrics, commit)
    SEEN["n"] += 1; SEEN["ts"] += int(not rej[0]); SEEN["sp"] += n_ok


def test_random_clusters_were_mostly_accepted():
    """(runs after the property test: the fast paths must have taken most of the drawn clusters, or the property proves little)"""
    assert SEEN["n"] >= 100 and SEEN["ts"] >= 0.5 * SEEN["n"] and SEEN["sp"] >= 0.8 * SEEN["n"], SEEN


@settings(max_examples=300, deadline=None)
@given(st.integers(0, 10**7), st.integers(1, 10**7))
def test_summary_f64_formatting_round_trips_with_shortest_digits(tp, fn):
    tot = np.zeros((abi.N_GROUPS, abi.N_METRICS), dtype=np.uint64)
    tot[0, abi.M_HAP:abi.M_HAP + 4] = [tp, fn, tp, fn]
    f = SummaryWriter("x", ["HAP"]).summary_text(tot).splitlines()[1].split("\t")
    recall = tp / (tp + fn)
    assert float(f[11]) == recall and float(f[12]) == recall                      # exact round trip
    digits = lambda s: s.replace(".", "").replace("-", "").lstrip("0").split("e")[0].rstrip("0")
    assert digits(f[11]) == digits(repr(recall))                                   # and no more digits than the shortest representation
    assert "e" not in f[11] or recall < 1e-5                                       # plain decimals down to 1e-5 (ryu's layout)


def test_skipped_noop_alt_is_not_taken_for_nothing_skipped():
    """Regression (found by the adversarial generator): a record whose ALT equals its REF has edit distance 0, so when it is
    skipped as incompatible -- it overlaps a spliced ALT -- the summed skip distance stays 0.  The zero-flip shortcut must not
    read that as "nothing skipped": optimize_gt_alleles drops the incompatible ALT child (exact_gt_optimizer.rs:293-305) and the
    variant is observed with one copy less.  Both scalar solvers against the oracle, and the expected labels spelled out."""
    import oracle_py as orc
    ref = b"A" * 80
    V = lambda p, a0, a1, vt: Variant(0, vt, p, a0, a1, max(len(a0), len(a1)))
    truth = [(V(14, b"AAA", b"AAA", VariantType.Indel), PhasedZygosity.HomozygousAlternate),
             (V(15, b"A", b"A", VariantType.Snv), PhasedZygosity.PhasedHet10),
             (V(15, b"A", b"G", VariantType.Snv), PhasedZygosity.UnphasedHeterozygous),
             (V(47, b"A", b"A", VariantType.Snv), PhasedZygosity.UnphasedHeterozygous)]
    query = [(V(50, b"A", b"AG", VariantType.Insertion), PhasedZygosity.PhasedHet10)]
    batch = _batch(ref, [truth, query])
    cfg = abi.CompareCfg(50, 0, 0, 0)
    cpu = orc.compare_batch(batch, [ref], cfg)
    assert list(cpu.var_expected[:5]) == [2, 1, 1, 1, 0] and list(cpu.var_observed[:5]) == [2, 0, 0, 1, 1]
    TS.check(batch, [ref], cfg)                         # (the thread solver may hand this one on: its pop budget)
    assert SP.check_solve(batch, [ref], cfg)[1] == 1
