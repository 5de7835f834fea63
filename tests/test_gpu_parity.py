"""GPU parity tests: the CUDA path, called through the C ABI, against (a) the reference's own
golden vectors and (b) the CPU oracle on seeded synthetic batches.  Bit-exact (integer work)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle_py as orc
from aardvark_b200 import abi, synth
from aardvark_b200.batch import CompareOutputs, RegionBatch, seq_offsets
from aardvark_b200.lib import Solver, compare_cfg, merge_cfg
from aardvark_b200.results import unpack_compare, unpack_merge
from aardvark_b200.types import CompareConfig, CompareRegion, MergeConfig
from golden_cases import COMPARE_CASES, ED_CASES, MERGE_CASES, MOCK_CHR1
from test_oracle_golden import check_compare_case

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def solver():
    s = Solver(0)
    yield s
    s.close()


def test_native_library_is_loaded(solver):
    maps = open("/proc/self/maps").read()
    assert "libaardvark_b200.so" in maps


def test_wfa_ed_golden(solver):  # dynamic_wfa.rs:289-403, sequence_alignment.rs:57-116
    ed = solver.wfa_ed_batch([(a, b) for a, b, _ in ED_CASES])
    assert list(ed) == [e for _, _, e in ED_CASES]


def test_wfa_ed_big_golden(solver):  # dynamic_wfa.rs:452-468: ED 5278
    g = json.load(open(os.path.join(HERE, "golden", "dwfa_big_early_termination.json")))
    c1, s23 = g["c1"].encode(), g["seq_23"].encode()
    ed = solver.wfa_ed_batch([(s23, c1), (c1, s23), (c1, c1)])
    assert list(ed) == [5278, 5278, 0]


def test_wfa_ed_random_vs_oracle(solver):
    rng = np.random.default_rng(7)
    pairs = []
    for _ in range(400):
        n = int(rng.integers(0, 300))
        a = synth.ACGT[rng.integers(0, 4, size=n)].copy()
        b = a.copy()
        for _ in range(int(rng.integers(0, 12))):
            if b.size == 0:
                break
            k = int(rng.integers(0, b.size))
            op = rng.integers(0, 3)
            if op == 0:
                b[k] = synth.ACGT[rng.integers(0, 4)]
            elif op == 1:
                b = np.delete(b, slice(k, k + int(rng.integers(1, 20))))
            else:
                b = np.insert(b, k, synth.ACGT[rng.integers(0, 4, size=int(rng.integers(1, 20)))])
        pairs.append((a.tobytes(), b.tobytes()))
    pairs.append((b"", b""))
    ed = solver.wfa_ed_batch(pairs)
    assert list(ed) == [orc.wfa_ed(a, b) for a, b in pairs]


@pytest.mark.parametrize("name,region,exp", COMPARE_CASES, ids=[c[0] for c in COMPARE_CASES])
def test_solve_compare_region_golden(solver, name, region, exp):  # waffle_solver.rs:896-1247
    solver.set_reference([MOCK_CHR1], ["mock_chr1"])
    bench = solver.solve_compare_region(region, CompareConfig())
    check_compare_case(bench, exp)


def _both_compare(solver, batch, contigs, cfg, **kw):
    gpu = solver.compare_batch(batch, cfg, **kw)
    cpu = orc.compare_batch(batch, contigs, compare_cfg(cfg), **kw)
    return gpu, cpu


def test_compare_golden_batch_vs_oracle(solver):
    """BASELINE config 1: the reference's unit-test clusters as one batch, all outputs incl. sequences."""
    solver.set_reference([MOCK_CHR1], ["mock_chr1"])
    regions = [CompareRegion(i, r.coordinates, r.truth_variants, r.truth_zygosity, r.query_variants, r.query_zygosity)
               for i, (_, r, _) in enumerate(COMPARE_CASES)]
    batch = RegionBatch.from_compare_regions(regions, solver.contig_index)
    off, plen = seq_offsets(batch)
    for shortcut in (False, True):
        cfg = CompareConfig(enable_sequences=True, enable_exact_shortcut=shortcut)
        gpu, cpu = _both_compare(solver, batch, [MOCK_CHR1], cfg, seq_off=off, seq_pool_len=plen)
        assert gpu.diff(cpu) == []


@pytest.mark.parametrize("name,region,cfg,expected", MERGE_CASES, ids=[c[0] for c in MERGE_CASES])
def test_solve_merge_region_golden(solver, name, region, cfg, expected):  # merge_solver.rs:242-348
    solver.set_reference([MOCK_CHR1], ["mock_chr1"])
    assert solver.solve_merge_region(region, cfg).merge_classification == expected


@pytest.mark.parametrize("seed", [20, 21, 22])
def test_compare_synthetic_vs_oracle(solver, seed):
    """chr20-shaped synthetic batch (configs[1] at 1/50 scale): every output array bit-exact."""
    ref, batch = synth.workload_chr20(scale=0.02, seed=seed)
    solver.set_reference([ref])
    strat_off = np.arange(batch.n_regions + 1, dtype=np.uint64)
    strat_idx = (np.arange(batch.n_regions) % 3).astype(np.uint32)
    kw = dict(strat_off=strat_off, strat_idx=strat_idx, n_strata=3)
    gpu, cpu = _both_compare(solver, batch, [ref], CompareConfig(enable_sequences=False), **kw)
    assert gpu.diff(cpu) == []
    assert int(gpu.solved_blocks[0]) == batch.n_regions
    # size-independent property: the reduction is the sum of the per-region rows
    assert (gpu.totals == gpu.region_metrics.sum(axis=0)).all()
    assert (gpu.strat_totals.sum(axis=0) == gpu.totals).all()


def test_compare_synthetic_sequences_and_branch_factor(solver):
    ref, batch = synth.workload_chr20(scale=0.005, seed=5)
    solver.set_reference([ref])
    off, plen = seq_offsets(batch)
    for mbf in (50, 2, 1):
        cfg = CompareConfig(enable_sequences=True, max_branch_factor=mbf)
        gpu, cpu = _both_compare(solver, batch, [ref], cfg, seq_off=off, seq_pool_len=plen)
        assert gpu.diff(cpu) == []


def test_compare_dense_clusters_vs_oracle(solver):
    """Dense het clusters (deep searches, workspace-tier escalation, auto-fail logic)."""
    p = synth.SynthParams(n_variants=3000, dense_frac=0.9, dense_mean=6.0, het_frac=0.95, phased_frac=0.2,
                          p_repr=0.05, p_gt_err=0.05, p_fn=0.05, p_fp=0.05)
    ref, batch = synth.workload_compare(60_000, p, seed=11)
    solver.set_reference([ref])
    gpu, cpu = _both_compare(solver, batch, [ref], CompareConfig(enable_sequences=False))
    assert gpu.diff(cpu) == []


def test_compare_sv_vs_oracle(solver):
    """SV / long-indel shaped batch (configs[3] scaled down): wide wavefronts, long windows."""
    p = synth.SynthParams(n_variants=300, sv_events=60, sv_max=1500, flank=1000)
    ref, batch = synth.workload_compare(400_000, p, seed=4)
    solver.set_reference([ref])
    gpu, cpu = _both_compare(solver, batch, [ref], CompareConfig(enable_sequences=False))
    assert gpu.diff(cpu) == []


def test_compare_large_sv_vs_oracle(solver):
    """Multi-kbp events: wavefronts thousands of diagonals wide, workspaces in the 2 MB / 64 MB global tiers."""
    p = synth.SynthParams(n_variants=40, sv_events=12, sv_min=1500, sv_max=5000, flank=1000)
    ref, batch = synth.workload_compare(300_000, p, seed=44)
    solver.set_reference([ref])
    gpu, cpu = _both_compare(solver, batch, [ref], CompareConfig(enable_sequences=False))
    assert gpu.diff(cpu) == []
    assert int(gpu.error_blocks[0]) == 0


def test_compare_edge_cases_vs_oracle(solver):
    """Empty batch, malformed regions, unsupported zygosity, non-ACGT bytes."""
    from aardvark_b200.types import Coordinates, PhasedZygosity as Z, Variant
    ref = b"ACGTNNNNacgtACGTRYKMACGTACGTACGTACGTACGT"
    solver.set_reference([ref], ["c"])
    empty = RegionBatch.from_compare_regions([], {"c": 0})
    out = solver.compare_batch(empty, CompareConfig(enable_sequences=False))
    assert int(out.solved_blocks[0]) == 0 and int(out.error_blocks[0]) == 0
    v = lambda p, a, b: Variant(0, abi.VT_SNV, p, a, b)
    regions = [
        CompareRegion(0, Coordinates("c", 0, 20), [v(5, b"N", b"n")], [Z.PhasedHet01], [v(5, b"N", b"n")], [Z.UnphasedHeterozygous]),
        CompareRegion(1, Coordinates("c", 0, 20), [v(9, b"c", b"C")], [Z.HomozygousAlternate], [v(9, b"c", b"G")], [Z.HomozygousAlternate]),
        CompareRegion(2, Coordinates("c", 0, 20), [v(5, b"N", b"A")], [Z.HomozygousReference], [], []),   # reference panics
        CompareRegion(3, Coordinates("c", 0, 20), [v(25, b"A", b"C")], [Z.PhasedHet01], [], []),           # outside window
        CompareRegion(4, Coordinates("c", 10, 400), [v(12, b"A", b"C")], [Z.PhasedHet01], [], []),         # window past contig
        CompareRegion(5, Coordinates("c", 0, 40), [], [], [v(30, b"A", b"C"), v(30, b"A", b"G"), v(30, b"A", b"T")],
                      [Z.UnphasedHeterozygous] * 3),
    ]
    batch = RegionBatch.from_compare_regions(regions, {"c": 0})
    gpu, cpu = _both_compare(solver, batch, [ref], CompareConfig(enable_sequences=False))
    assert gpu.diff(cpu) == []
    assert list(gpu.status[:5]) == [0, 0, abi.ST_BAD_ZYGOSITY, abi.ST_BAD_INPUT, abi.ST_BAD_INPUT]
    assert int(gpu.error_blocks[0]) == 3


def test_compare_single_snv_pairs_closed_form_vs_oracle(solver):
    """One truth + one query SNV: every zygosity pair, equal / different ALT, ALT == reference base, allele0 that
    disagrees with the reference, odd type labels and raw_allele_space 0 -- the shapes around k_compare_simple's
    closed form -- under branch factors 1..50, the exact shortcut and the sequence bundle."""
    import itertools
    from aardvark_b200.types import Coordinates, PhasedZygosity as Z, Variant, VariantType
    rng = np.random.default_rng(11)
    ref = bytes(synth.ACGT[rng.integers(0, 4, size=4000)])
    solver.set_reference([ref], ["c"])
    zygs = [Z.Unknown, Z.HomozygousReference, Z.UnphasedHeterozygous, Z.PhasedHet01, Z.PhasedHet10, Z.HomozygousAlternate]
    regions, zero_raw = [], []
    rid = 0
    for zt, zq in itertools.product(zygs, zygs):
        for shape in range(13):
            p = 60 + 7 * (rid % 500)
            r0 = ref[p:p + 1]
            alts = [b for b in (b"A", b"C", b"G", b"T") if b != r0]
            ta, qa, t0, q0, tp, qp = alts[0], alts[0], r0, r0, p, p
            tt = qt = VariantType.Snv
            if shape == 1: qa = alts[1]            # different ALT
            if shape == 2: ta = qa = r0            # ALT == reference base
            if shape == 3: t0 = q0 = alts[2]       # allele0 disagrees with the reference
            if shape == 4: qp = p + 1; q0 = ref[qp:qp + 1]; qa = [b for b in (b"A", b"C", b"G", b"T") if b != q0][0]
            if shape == 5: qt = VariantType.Indel  # odd label on a 1 -> 1 substitution
            if shape == 6: zero_raw.append(rid)
            if shape >= 7:                          # indel pairs around the closed form for anchored pure indels
                ins = bytes(synth.ACGT[rng.integers(0, 4, size=int(rng.integers(1, 9)))])
                dl = int(rng.integers(2, 10))
                if shape == 7: tt = qt = VariantType.Insertion; t0 = q0 = r0; ta = qa = r0 + ins
                if shape == 8: tt = qt = VariantType.Deletion; t0 = q0 = ref[p:p + dl]; ta = qa = r0
                if shape == 9: tt = qt = VariantType.Insertion; t0 = q0 = r0; ta = qa = alts[0] + ins      # anchor != reference
                if shape == 10: tt = qt = VariantType.Insertion; t0 = q0 = r0; ta = r0 + ins; qa = r0 + ins[::-1] + b"A"
                if shape == 11: tt = qt = VariantType.Indel; t0 = q0 = r0; ta = qa = r0 + ins              # odd but supported label
                if shape == 12: tt = VariantType.Deletion; qt = VariantType.Indel; t0 = q0 = ref[p:p + dl]; ta = qa = r0
            regions.append(CompareRegion(rid, Coordinates("c", p - 50, p + 52), [Variant(0, tt, tp, t0, ta)], [zt],
                                         [Variant(0, qt, qp, q0, qa)], [zq]))
            rid += 1
    batch = RegionBatch.from_compare_regions(regions, {"c": 0})
    for r in zero_raw:   # raw_allele_space 0 on the truth record: the record-basepair underflow check must still fire
        batch.raw_allele_space[int(batch.var_off[2 * r])] = 0
    off, plen = seq_offsets(batch)
    for mbf in (1, 2, 3, 4, 50):
        gpu, cpu = _both_compare(solver, batch, [ref], CompareConfig(enable_sequences=False, max_branch_factor=mbf))
        assert gpu.diff(cpu) == [], f"max_branch_factor={mbf}"
    gpu, cpu = _both_compare(solver, batch, [ref], CompareConfig(enable_sequences=False, enable_exact_shortcut=True))
    assert gpu.diff(cpu) == []
    gpu, cpu = _both_compare(solver, batch, [ref], CompareConfig(enable_sequences=True), seq_off=off, seq_pool_len=plen)
    assert gpu.diff(cpu) == []


def test_compare_multi_contig_vs_oracle(solver):
    """BASELINE configs[2] shape: one batch over several contigs of different lengths (contig index -> reference)."""
    parts = [synth.workload_chr20(scale=sc, seed=sd) for sc, sd in ((0.003, 101), (0.0012, 102), (0.002, 103))]
    contigs = [p[0] for p in parts]
    batch = RegionBatch.concat([p[1] for p in parts])
    assert sorted(set(batch.contig.tolist())) == [0, 1, 2]
    solver.set_reference(contigs)
    gpu, cpu = _both_compare(solver, batch, contigs, CompareConfig(enable_sequences=False))
    assert gpu.diff(cpu) == []
    assert int(gpu.error_blocks[0]) == 0
    # a contig bin solved on its own gives the same rows (sharding property)
    lo = parts[0][1].n_regions
    hi = lo + parts[1][1].n_regions
    part = solver.compare_batch(batch.slice_regions(lo, hi), CompareConfig(enable_sequences=False))
    assert np.array_equal(part.region_metrics[:hi - lo], gpu.region_metrics[lo:hi])
    assert np.array_equal(part.status[:hi - lo], gpu.status[lo:hi])


def _same_batch(a: RegionBatch, b: RegionBatch):
    for f in ("region_id", "contig", "start", "end", "var_off", "position", "variant_type", "zygosity", "raw_allele_space",
              "allele_off", "a0_len", "a1_len"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    n = int(a.a0_len.sum() + a.a1_len.sum())
    assert np.array_equal(a.allele_pool[:n], b.allele_pool[:n])
    assert a.n_inputs == b.n_inputs


def test_build_regions_vs_host_builder(solver):
    """Device region builder (region_generation.rs:352-469) against the host builder the generator uses: compare and
    merge call sets, flanks 0 / 50 / 1000, ties across inputs, an empty input, variants reaching past the contig end;
    then the batch is solved where it was built (no upload) and must equal the uploaded host-built batch."""
    from aardvark_b200.batch import CallSets
    ref, inputs = synth.callsets_compare(400_000, synth.SynthParams(n_variants=900), seed=71)
    solver.set_reference([ref])
    for flank in (0, 50, 1000):
        host = synth.cluster_regions(inputs, len(ref), flank)
        dev = solver.build_regions(CallSets(inputs), 0, flank)
        _same_batch(dev, host)
    host = synth.cluster_regions(inputs, len(ref), 50)
    solver.build_regions(CallSets(inputs), 0, 50, download=False)
    solver.run_resident(CompareConfig(enable_sequences=False))
    built = solver.download(CompareOutputs(host))
    assert built.diff(solver.compare_batch(host, CompareConfig(enable_sequences=False))) == []
    # K = 5 call sets with dropout (merge inputs), first_region_id offset
    ref5, sets, flank5 = synth.callsets_merge(200_000, 500, n_sets=5, seed=72)
    solver.set_reference([ref5])
    host = synth.cluster_regions(sets, len(ref5), flank5, first_region_id=1000)
    _same_batch(solver.build_regions(CallSets(sets), 0, flank5, first_region_id=1000), host)
    # hand-made corner cases on a short contig
    small = bytes(synth.ACGT[np.random.default_rng(5).integers(0, 4, size=600)])
    solver.set_reference([small], ["s"])
    rec = lambda p, a0, a1: synth.make_rec(p, a0, a1, abi.ZYG_HOM_ALT)
    A = [rec(10, small[10:11], b"T" if small[10:11] != b"T" else b"G"), rec(10, small[10:11], b"AAAC"), rec(300, small[300:303], small[300:301]),
         rec(598, small[598:600], b"A"), rec(599, small[599:600] + b"A", b"C")]          # the last one reaches past the end: dropped
    B = [rec(10, small[10:11], b"C" if small[10:11] != b"C" else b"G"), rec(61, small[61:62], b"A" if small[61:62] != b"A" else b"C"),
         rec(352, small[352:353], b"T" if small[352:353] != b"T" else b"C")]
    for sets3, flank in (([A, B], 50), ([A, [], B], 50), ([[], []], 50), ([B, A], 0), ([A, B], 5000)):
        _same_batch(solver.build_regions(CallSets(sets3), 0, flank), synth.cluster_regions(sets3, len(small), flank))


def test_compare_two_snv_pairs_closed_form_vs_oracle(solver):
    """Two truth + two query SNVs: every zygosity combination of the four records (allele-bearing codes), at distances
    1 / 2 / 30, plus the neighbouring shapes that must take the general path (different ALT, ALT == reference base,
    same position, an insertion instead of a substitution, a third query record) -- around k_compare_simple's
    two-pair closed form -- under branch factors 4, 15, 16, 50."""
    import itertools
    from aardvark_b200.types import Coordinates, PhasedZygosity as Z, Variant, VariantType
    rng = np.random.default_rng(12)
    ref = bytes(synth.ACGT[rng.integers(0, 4, size=6000)])
    solver.set_reference([ref], ["c"])
    zygs = [Z.UnphasedHeterozygous, Z.PhasedHet01, Z.PhasedHet10, Z.HomozygousAlternate]
    other = lambda b_, k=0: [x for x in (b"A", b"C", b"G", b"T") if x != b_][k]
    snv = lambda p, alt: Variant(0, VariantType.Snv, p, ref[p:p + 1], alt)
    regions = []
    rid = 0
    for z in itertools.product(zygs, repeat=4):
        for shape in range(9):
            p = 100 + 11 * (rid % 450)
            d = (1, 2, 30)[rid % 3]
            q = p + d
            a, b2 = other(ref[p:p + 1]), other(ref[q:q + 1])
            tv, qv = [snv(p, a), snv(q, b2)], [snv(p, a), snv(q, b2)]
            tz, qz = [z[0], z[1]], [z[2], z[3]]
            if shape == 1: qv[1] = snv(q, other(ref[q:q + 1], 1))                       # different ALT on the second pair
            if shape == 2: tv[0] = qv[0] = snv(p, ref[p:p + 1])                         # ALT == reference base
            if shape == 3: tv[1] = qv[1] = snv(p, other(ref[p:p + 1], 1))               # both at the same position
            if shape == 4: tv[1] = qv[1] = Variant(0, VariantType.Insertion, q, ref[q:q + 1], ref[q:q + 1] + b"GT")
            if shape == 5: qv.append(snv(q + 3, other(ref[q + 3:q + 4]))); qz.append(Z.HomozygousAlternate)
            if shape == 6: qz = [z[2], Z.HomozygousAlternate if z[1] != Z.HomozygousAlternate else Z.PhasedHet01]   # copies differ
            if shape == 7: tv[0] = Variant(0, VariantType.Snv, p, other(ref[p:p + 1], 2), a); qv[0] = tv[0]        # allele0 != reference
            regions.append(CompareRegion(rid, Coordinates("c", p - 50, q + 55), tv, tz, qv, qz))
            rid += 1
    batch = RegionBatch.from_compare_regions(regions, {"c": 0})
    for mbf in (4, 15, 16, 50):
        gpu, cpu = _both_compare(solver, batch, [ref], CompareConfig(enable_sequences=False, max_branch_factor=mbf))
        assert gpu.diff(cpu) == [], f"max_branch_factor={mbf}"
    gpu, cpu = _both_compare(solver, batch, [ref], CompareConfig(enable_sequences=False, enable_exact_shortcut=True))
    assert gpu.diff(cpu) == []


def test_merge_synthetic_vs_oracle(solver):
    ref, batch = synth.workload_merge(150_000, 400, n_sets=5, seed=38)
    solver.set_reference([ref])
    for cfg in (MergeConfig(majority_voting_enabled=True), MergeConfig(no_conflict_enabled=True),
                MergeConfig(conflict_selection=2), MergeConfig()):
        gpu = solver.merge_batch(batch, cfg)
        cpu = orc.merge_batch(batch, [ref], merge_cfg(cfg))
        assert gpu.diff(cpu) == []


def test_resident_mode_matches_batch_mode(solver):
    ref, batch = synth.workload_chr20(scale=0.004, seed=3)
    solver.set_reference([ref])
    a = solver.compare_batch(batch, CompareConfig(enable_sequences=False))
    solver.upload(batch)
    solver.run_resident(CompareConfig(enable_sequences=False))
    solver.run_resident(CompareConfig(enable_sequences=False))   # idempotent
    b = solver.download(CompareOutputs(batch))
    assert a.diff(b) == []
    assert solver.launch_count() > 0
    w = solver.last_work()
    assert w["search_pops"] > 0 and w["alignments"] > 0
