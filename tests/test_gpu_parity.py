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
    solver.run_resident(CompareConfig(enable_sequences=False), region_metrics=True)
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
    solver.run_resident(CompareConfig(enable_sequences=False), region_metrics=True)
    solver.run_resident(CompareConfig(enable_sequences=False), region_metrics=True)   # idempotent
    b = solver.download(CompareOutputs(batch))
    assert a.diff(b) == []
    # without per-region rows: the summary counters come from the in-kernel totals alone and must be the same
    solver.run_resident(CompareConfig(enable_sequences=False))
    c = solver.download(CompareOutputs(batch, region_metrics=False))
    assert a.diff(c) == []
    with pytest.raises(Exception):
        solver.download(CompareOutputs(batch))               # rows were not kept by that run
    assert solver.launch_count() > 0
    w = solver.last_work()
    assert w["search_pops"] > 0 and w["alignments"] > 0


# ------------------------------------------------------------------------------------------------------------------
# Round-2 parity tests: the regimes the first round's driver-run suite did not reach

def _sv_corner_batch(seed=91):
    """Hand-made SV clusters (min gap 1000, region_generation.rs:621-626 admits alleles up to 10 kbp): single events of
    6 / 8 / 10 kbp as INS and DEL (identical on both sides, hom and het), a sequence-divergent 6 kbp INS, a right-shifted
    representation of a tandem INS, a 3 kbp DEL against a 2 kbp INS at the same place (last row),
    an FN and an FP event."""
    rng = np.random.default_rng(seed)
    ref = synth.random_reference(260_000, rng)
    T, Q = [], []
    rs = lambda n: synth._rand_seq(rng, n)
    pos = 3000

    def ins(p, n, tandem=False):
        anchor = bytes([int(ref[p])])
        body = ref[p + 1:p + 1 + n].tobytes() if tandem else rs(n)
        return p, anchor, anchor + body

    def dele(p, n):
        return p, ref[p:p + 1 + n].tobytes(), bytes([int(ref[p])])

    HOM, HET, H01 = abi.ZYG_HOM_ALT, abi.ZYG_UNPHASED_HET, abi.ZYG_PHASED_HET01
    for n, zt, zq in ((6000, HOM, HOM), (8000, HET, H01), (10000, H01, HET)):
        p, a0, a1 = ins(pos, n); T.append(synth.make_rec(p, a0, a1, zt, sv=True)); Q.append(synth.make_rec(p, a0, a1, zq, sv=True)); pos += 14000
        p, a0, a1 = dele(pos, n); T.append(synth.make_rec(p, a0, a1, zt, sv=True)); Q.append(synth.make_rec(p, a0, a1, zq, sv=True)); pos += 14000 + n
    # divergent insertion: ~2 % substitutions inside the inserted sequence
    p, a0, a1 = ins(pos, 6000)
    arr = np.frombuffer(a1, dtype=np.uint8).copy()
    idx = rng.integers(1, arr.size, size=120)
    arr[idx] = synth.ACGT[rng.integers(0, 4, size=120)]
    T.append(synth.make_rec(p, a0, a1, HOM, sv=True)); Q.append(synth.make_rec(p, a0, arr.tobytes(), HOM, sv=True)); pos += 14000
    # tandem duplication, query right-shifted inside the repeat (same haplotype, other record)
    p, a0, a1 = ins(pos, 7000, tandem=True)
    sh = synth._shift_insertion(ref, p, a0, a1, rng)
    T.append(synth.make_rec(p, a0, a1, HET, sv=True)); Q.append(synth.make_rec(sh[0], sh[1], sh[2], HET, sv=True)); pos += 16000
    # 3 kbp deletion vs 2 kbp insertion at the same position
    p, a0, a1 = dele(pos, 3000); T.append(synth.make_rec(p, a0, a1, HOM, sv=True))
    p, a0, a1 = ins(pos, 2000); Q.append(synth.make_rec(p, a0, a1, HOM, sv=True)); pos += 12000
    # false negative (9 kbp DEL missing from the query) and false positive (6.5 kbp INS only in the query)
    p, a0, a1 = dele(pos, 9000); T.append(synth.make_rec(p, a0, a1, HET, sv=True)); pos += 20000
    p, a0, a1 = ins(pos, 6500); Q.append(synth.make_rec(p, a0, a1, HOM, sv=True)); pos += 12000
    assert pos < ref.size - 12000
    return ref, synth.cluster_regions([T, Q], ref.size, 1000)


def test_compare_sv_6_to_10kbp_vs_oracle(solver):
    """configs[3] goes to 10 kbp events: edit distances up to ~10^4, wavefronts 2 * 10^4 wide (dynamic_wfa.rs:452-468 is
    the reference's own wide vector, ED 5278)."""
    ref, batch = _sv_corner_batch()
    assert int(np.maximum(batch.a0_len, batch.a1_len).max()) == 10001
    solver.set_reference([ref])
    gpu, cpu = _both_compare(solver, batch, [ref], CompareConfig(enable_sequences=False))
    assert gpu.diff(cpu) == []
    assert int(gpu.error_blocks[0]) == 0 and int(gpu.solved_blocks[0]) == batch.n_regions
    assert int(max(gpu.ed1.max(), gpu.ed2.max())) >= 9000          # the FN deletion: ED(truth hap, query hap) ~ its length


def _solver_with_env(**env):
    """A context of its own created under test knobs (read once in avk_create)."""
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return Solver(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_compare_forced_last_resort_tiers_vs_oracle():
    """The 2 GB cooperative tier and the DWFA_COOP_SPILL fallback (a wavefront that outgrows shared memory is finished on
    the warp path) are reached by a fraction of a per cent of configs[3]; test knobs shrink the tiers so that a small SV
    batch drives every cluster with an edit-distance bound >= 32 through them."""
    p = synth.SynthParams(n_variants=120, sv_events=30, sv_min=60, sv_max=900, flank=1000)
    ref, batch = synth.workload_compare(250_000, p, seed=45)
    cfg = CompareConfig(enable_sequences=False)
    cpu = orc.compare_batch(batch, [ref], compare_cfg(cfg))
    # (1) tiny first cooperative tier: clusters overflow it and finish in the last-resort tier
    # (2) tiny shared-memory wavefront: the CTA-wide DWFA spills back to the warp path mid-alignment
    for env in (dict(AVK_TEST_WIDE_B0=32, AVK_TEST_COOP_ARENA_MB=1, AVK_TEST_COOP_ARENA1_MB=256),
                dict(AVK_TEST_WIDE_B0=32, AVK_TEST_COOP_CAP_INTS=200)):
        s = _solver_with_env(**env)
        try:
            s.set_reference([ref])
            gpu = s.compare_batch(batch, cfg)
            assert gpu.diff(cpu) == [], env
            assert s.last_tier_overflow()[2] > 0, "no cluster reached the cooperative tiers"
        finally:
            s.close()


def test_merge_k5_tenth_chr20_vs_oracle(solver):
    """configs[4] shape: 5 call sets with per-set error rates and 5 % dropout, majority voting, 0.1 x chr20."""
    ref, batch = synth.workload_merge(int(synth.CHR20_LEN * 0.1), 15_000, n_sets=5, seed=38)
    solver.set_reference([ref])
    cfg = MergeConfig(majority_voting_enabled=True)
    gpu = solver.merge_batch(batch, cfg)
    cpu = orc.merge_batch(batch, [ref], merge_cfg(cfg), n_threads=orc.num_threads())
    assert gpu.diff(cpu) == []
    assert batch.n_regions > 10_000 and len(set(gpu.classification.tolist())) >= 3


def test_wgs_24_contigs_and_8_bin_sharding_vs_oracle(solver):
    """configs[2] at 1/20 scale: 24 contigs in one batch, bit-exact against the oracle; then the multi-GPU property --
    8 contiguous bins solved separately (avk_compare_batch_range writes each bin at its offset of ONE output set)
    reproduce the whole batch, and the bins' counters add up to its totals."""
    from aardvark_b200.dist import partition_regions
    refs, batch = synth.workload_wgs(scale=0.05, seed=38, workers=1)
    assert len(refs) == 24 and batch.n_regions > 150_000
    solver.set_reference(refs)
    cfg = CompareConfig(enable_sequences=False)
    gpu, cpu = _both_compare(solver, batch, refs, cfg)
    assert gpu.diff(cpu) == []
    bins = partition_regions(batch, 8)
    assert bins[0][0] == 0 and bins[-1][1] == batch.n_regions and all(a[1] == b[0] for a, b in zip(bins, bins[1:]))
    sharded = CompareOutputs(batch)
    tot = np.zeros_like(gpu.totals)
    solved = 0
    for lo, hi in bins:
        solver.compare_batch_range(batch, lo, hi, cfg, out=sharded)
        tot += sharded.totals
        solved += int(sharded.solved_blocks[0])
    sharded.totals[...] = tot
    sharded.solved_blocks[0] = solved
    sharded.totals_mask[0] = gpu.totals_mask[0]
    assert sharded.diff(gpu) == []


def test_multi_context_entry_point_vs_single(solver):
    """avk_compare_batch_multi / avk_merge_batch_multi: one host thread per context inside the library, contiguous bins,
    results written at the bins' offsets, counters added on the host.  Uses two GPUs when the box has them, else two
    contexts on GPU 0 (the same code path: a context per bin)."""
    import torch
    from aardvark_b200.lib import MultiSolver
    devs = [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]
    ref, batch = synth.workload_chr20(scale=0.02, seed=23)
    solver.set_reference([ref])
    cfg = CompareConfig(enable_sequences=False)
    strat_off = np.arange(batch.n_regions + 1, dtype=np.uint64)
    strat_idx = (np.arange(batch.n_regions) % 3).astype(np.uint32)
    one = solver.compare_batch(batch, cfg, strat_off=strat_off, strat_idx=strat_idx, n_strata=3)
    one_tot = solver.compare_batch(batch, cfg, region_metrics=False)
    m = MultiSolver(devs + devs[:1])          # three bins
    try:
        m.set_reference([ref])
        multi = m.compare_batch(batch, cfg, strat_off=strat_off, strat_idx=strat_idx, n_strata=3)
        assert multi.diff(one) == []
        assert m.compare_batch(batch, cfg, region_metrics=False).diff(one_tot) == []
        refm, mb = synth.workload_merge(200_000, 600, n_sets=5, seed=9)
        solver.set_reference([refm])
        m.set_reference([refm])
        mc = MergeConfig(majority_voting_enabled=True)
        assert m.merge_batch(mb, mc).diff(solver.merge_batch(mb, mc)) == []
    finally:
        m.close()


def test_wfa_ed_100k_random_pairs_vs_oracle(solver):
    """10^5 random pairs, lengths up to 2 kbp (SURVEY section 7's minimum slice): substitutions, insertions, deletions and
    block moves; GPU batched wfa_ed against the oracle's (sequence_alignment.rs:9-13)."""
    rng = np.random.default_rng(2024)
    n_pairs = 100_000
    lens = np.minimum(2000, (rng.exponential(220.0, n_pairs)).astype(np.int64))
    lens[::97] = 2000
    lens[::101] = 0
    pairs = []
    for i in range(n_pairs):
        n = int(lens[i])
        a = synth.ACGT[rng.integers(0, 4, size=n)]
        b = a
        k = int(rng.integers(0, 9))
        if k and n:
            b = a.copy()
            where = rng.integers(0, n, size=k)
            b[where] = synth.ACGT[rng.integers(0, 4, size=k)]
            if i % 3 == 0:                                    # an indel block
                c = int(rng.integers(0, n)); ln = int(rng.integers(1, 40))
                b = np.concatenate([b[:c], b[c + ln:]]) if i % 2 else np.concatenate([b[:c], synth.ACGT[rng.integers(0, 4, size=ln)], b[c:]])
        pairs.append((a.tobytes(), b.tobytes()))
    gpu = solver.wfa_ed_batch(pairs)
    cpu = np.array([orc.wfa_ed(a, b) for a, b in pairs], dtype=np.uint32)
    assert np.array_equal(gpu, cpu)
    assert int(cpu.max()) > 30 and int((cpu == 0).sum()) > 1000


def test_pipelined_single_gpu_call_vs_unpipelined(solver):
    """avk_compare_batch on a large batch runs as a pipeline over sibling contexts of the same GPU (bins' uploads, kernels and
    downloads overlap); the result must not depend on it.  The knob lowers the size threshold so that a small batch takes the
    pipelined path."""
    ref, batch = synth.workload_chr20(scale=0.03, seed=29)
    cfg = CompareConfig(enable_sequences=False)
    solver.set_reference([ref])
    strat_off = np.arange(batch.n_regions + 1, dtype=np.uint64)
    strat_idx = (np.arange(batch.n_regions) % 2).astype(np.uint32)
    plain = solver.compare_batch(batch, cfg, strat_off=strat_off, strat_idx=strat_idx, n_strata=2)
    plain_tot = solver.compare_batch(batch, cfg, region_metrics=False)
    s = _solver_with_env(AVK_PIPELINE_MIN_REGIONS=100, AVK_PIPELINE_BINS=7)
    try:
        s.set_reference([ref])
        assert s.compare_batch(batch, cfg, strat_off=strat_off, strat_idx=strat_idx, n_strata=2).diff(plain) == []
        assert s.compare_batch(batch, cfg, region_metrics=False).diff(plain_tot) == []
        off, plen = seq_offsets(batch)
        cs = CompareConfig(enable_sequences=True)
        assert s.compare_batch(batch, cs, seq_off=off, seq_pool_len=plen).diff(solver.compare_batch(batch, cs, seq_off=off, seq_pool_len=plen)) == []
    finally:
        s.close()


def test_concurrent_lanes_single_gpu_call_vs_one_piece(solver):
    """avk_compare_batch on a large batch cuts it into contiguous bins that sibling contexts of the same GPU solve at the same
    time (own streams, the owner's reference and stratification tables); the result must not depend on it.  The knobs lower
    the size threshold so that a small batch takes that path, with host-side strata, device-side strata, totals only, and
    result sequences."""
    from aardvark_b200.batch import StratIntervals
    ref, batch = synth.workload_chr20(scale=0.03, seed=31)
    cfg = CompareConfig(enable_sequences=False)
    solver.set_reference([ref])
    strat_off = np.arange(batch.n_regions + 1, dtype=np.uint64)
    strat_idx = (np.arange(batch.n_regions) % 2).astype(np.uint32)
    plain = solver.compare_batch(batch, cfg, strat_off=strat_off, strat_idx=strat_idx, n_strata=2)
    plain_tot = solver.compare_batch(batch, cfg, region_metrics=False)
    L = len(ref)
    strat = StratIntervals([[(0, 0, L // 3), (0, L // 2, L)], [(0, L // 4, 3 * L // 4)]], 1)
    solver.set_stratifications(strat)
    plain_dev = solver.compare_batch(batch, cfg, n_strata=2, device_strata=True, containment=True)
    for lanes in (2, 5):
        s = _solver_with_env(AVK_LANE_MIN_REGIONS=100, AVK_LANES=lanes)
        try:
            s.set_reference([ref])
            assert s.compare_batch(batch, cfg, strat_off=strat_off, strat_idx=strat_idx, n_strata=2).diff(plain) == []
            assert s.compare_batch(batch, cfg, region_metrics=False).diff(plain_tot) == []
            s.set_stratifications(strat)
            assert s.compare_batch(batch, cfg, n_strata=2, device_strata=True, containment=True).diff(plain_dev) == []
            off, plen = seq_offsets(batch)
            cs = CompareConfig(enable_sequences=True)
            assert s.compare_batch(batch, cs, seq_off=off, seq_pool_len=plen).diff(solver.compare_batch(batch, cs, seq_off=off, seq_pool_len=plen)) == []
            with pytest.raises(Exception):
                s.run_resident(cfg)        # a lane call leaves nothing resident: loud, not one bin's worth of results
        finally:
            s.close()


def test_solver_pool_batches_in_flight_vs_oracle():
    """Several batches in flight on one GPU (avk_create_lane: lanes share the owner's reference and stratification tables, one
    host thread per context): every batch's result equals the oracle's, whatever was running beside it; a lane refuses its
    own reference; the resident entry points work on a lane."""
    from aardvark_b200.batch import StratIntervals
    from aardvark_b200.lib import SolverPool
    ref, batch = synth.workload_chr20(scale=0.04, seed=37)
    cfg = CompareConfig(enable_sequences=False)
    n = batch.n_regions
    cuts = [0, n // 7, n // 3, n // 2, (3 * n) // 4, n]
    parts = [batch.slice_regions(cuts[i], cuts[i + 1]) for i in range(5)] * 3        # 15 calls over 4 contexts
    L = len(ref)
    strat = StratIntervals([[(0, 0, L // 3), (0, L // 2, L)], [(0, L // 4, 3 * L // 4)]], 1)
    pool = SolverPool(0, 4)
    try:
        pool.set_reference([ref])
        pool.set_stratifications(strat)
        got = pool.compare_batches(parts, cfg, n_strata=2, device_strata=True, containment=True)
        want = [None] * 5
        for i, (b, g) in enumerate(zip(parts, got)):
            if want[i % 5] is None:
                o = orc.compare_batch(b, [ref], compare_cfg(cfg))
                masks = orc.containments(b, strat)
                want[i % 5] = (o, masks)
            o, masks = want[i % 5]
            assert g.diff(o) == [], i
            assert np.array_equal(g.containment[:b.n_regions], masks)
        lane = pool.solvers[1]
        with pytest.raises(Exception):
            lane.set_reference([ref])
        lane.upload(parts[2])
        lane.run_resident(cfg)
        o = CompareOutputs(parts[2], region_metrics=False)
        lane.download(o)
        assert np.array_equal(o.ed1, want[2][0].ed1) and np.array_equal(o.var_class, want[2][0].var_class)
    finally:
        pool.close()


def test_config0_golden_clusters_with_stratification(solver):
    """BASELINE configs[0] (the bundle does not exist: SURVEY 8d substitutes the unit-test clusters on contigs `mock` /
    `mock2` with test_data/example_stratification): the golden clusters, shifted onto both contigs, are solved with the
    DEVICE containment lookup (avk_set_stratifications); containment masks and stratified sums against the oracle's
    restatement of Stratifications::containments + var_coordinates()."""
    from aardvark_b200.batch import StratIntervals, masks_to_membership
    from aardvark_b200.types import Coordinates
    from test_strat import EXAMPLE1, EXAMPLE2
    mock = MOCK_CHR1 * 2                      # 50 bp: the golden clusters live in [0, 25), a second copy in [25, 50)
    regions = []
    for ci, chrom in enumerate(("mock", "mock2")):
        for shift in (0, 25):
            for (_, r, _) in COMPARE_CASES:
                sh = lambda vs: [type(v)(v.vcf_index, v.variant_type, v.position + shift, v.allele0, v.allele1) for v in vs]
                regions.append(CompareRegion(len(regions), Coordinates(chrom, r.coordinates.start + shift, r.coordinates.end + shift),
                                             sh(r.truth_variants), r.truth_zygosity, sh(r.query_variants), r.query_zygosity))
    batch = RegionBatch.from_compare_regions(regions, {"mock": 0, "mock2": 1})
    strat = StratIntervals([EXAMPLE1, EXAMPLE2], n_contigs=2)
    solver.set_reference([mock, mock], ["mock", "mock2"])
    solver.set_stratifications(strat)
    cfg = CompareConfig(enable_sequences=False)
    gpu = solver.compare_batch(batch, cfg, n_strata=2, device_strata=True, containment=True)
    masks = orc.containments(batch, strat)
    assert np.array_equal(gpu.containment[:batch.n_regions], masks)
    assert len(set(masks.tolist())) >= 3                      # regions inside neither, one and both strata
    so, si = masks_to_membership(masks)
    cpu = orc.compare_batch(batch, [mock, mock], compare_cfg(cfg), strat_off=so, strat_idx=si, n_strata=2)
    assert gpu.diff(cpu) == []
    assert int(gpu.solved_blocks[0]) == batch.n_regions
    # the caller-provided membership list gives the same stratified sums
    assert solver.compare_batch(batch, cfg, strat_off=so, strat_idx=si, n_strata=2).diff(cpu) == []


def test_cpp_abi_harness():
    """A C++ program (tests/abi_harness.cpp) calls the C ABI directly -- the reference's run_compare around one batched call
    (src/main.rs:217-279) -- and prints what the reference's first solve_compare_region test asserts (waffle_solver.rs:897-930)
    plus a false negative."""
    import subprocess
    root = os.path.dirname(HERE)
    exe = os.path.join(HERE, "_build", "abi_harness")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    libdir = os.path.join(root, "aardvark_b200", "csrc")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-o", exe, os.path.join(HERE, "abi_harness.cpp"),
                           "-L" + libdir, "-laardvark_b200", "-Wl,-rpath," + libdir])
    txt = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()
    assert txt[0] == "status 0 0 ed 0 0 1 1 solved 2 errors 0"
    assert txt[1] == "variants 1/1/1 1/1/1 2/0/2"                    # TP, TP, FN (expected 2, observed 0)
    assert txt[2] == "joint GT tp 1 fn 1 qtp 1 qfp 0 BASEPAIR tp 2 fn 4 qtp 2 qfp 0"


def test_warp_only_pipeline_vs_oracle():
    """The pipeline without the thread-per-cluster stage (what small batches get by default): warp search / score / fused /
    team kernels only."""
    s = _solver_with_env(AVK_NO_THREAD_STAGE=1)
    try:
        for seed in (24, 25):
            ref, batch = synth.workload_chr20(scale=0.02, seed=seed)
            s.set_reference([ref])
            cfg = CompareConfig(enable_sequences=False)
            gpu = s.compare_batch(batch, cfg)
            cpu = orc.compare_batch(batch, [ref], compare_cfg(cfg))
            assert gpu.diff(cpu) == []
        p = synth.SynthParams(n_variants=2000, dense_frac=0.9, dense_mean=6.0, het_frac=0.95, phased_frac=0.2, p_repr=0.05, p_gt_err=0.05, p_fn=0.05, p_fp=0.05)
        ref, batch = synth.workload_compare(40_000, p, seed=12)
        s.set_reference([ref])
        assert s.compare_batch(batch, CompareConfig(enable_sequences=False)).diff(orc.compare_batch(batch, [ref], compare_cfg(CompareConfig(enable_sequences=False)))) == []
    finally:
        s.close()


def test_exact_gt_expansion_cap_timeout(solver):
    """AVK_ST_TIMEOUT, the deterministic stand-in for the reference's 300 s bail-out (exact_gt_optimizer.rs:165,174-176): a cap
    of 3 expansions per optimize_gt_alleles call trips on dense het clusters.  The device prunes exact-GT searches the oracle
    runs in full, so it can only time out where the oracle does; wherever both agree on the status every output
    is identical, and with the default cap nothing times out."""
    p = synth.SynthParams(n_variants=1500, dense_frac=0.9, dense_mean=6.0, het_frac=0.95, phased_frac=0.2,
                          p_repr=0.05, p_gt_err=0.08, p_fn=0.08, p_fp=0.08)
    ref, batch = synth.workload_compare(30_000, p, seed=17)
    solver.set_reference([ref])
    cfg = CompareConfig(enable_sequences=False, exact_gt_max_expansions=3)
    gpu = solver.compare_batch(batch, cfg)
    cpu = orc.compare_batch(batch, [ref], compare_cfg(cfg))
    n = batch.n_regions
    g_to, c_to = gpu.status[:n] == abi.ST_TIMEOUT, cpu.status[:n] == abi.ST_TIMEOUT
    assert g_to.sum() > 0 and c_to[g_to].all()
    same = gpu.status[:n] == cpu.status[:n]
    assert same.mean() > 0.9
    assert (gpu.region_metrics[:n][same] == cpu.region_metrics[:n][same]).all()
    assert (gpu.ed1[:n][same] == cpu.ed1[:n][same]).all()
    dflt = solver.compare_batch(batch, CompareConfig(enable_sequences=False))
    assert (dflt.status[:n] != abi.ST_TIMEOUT).all()
    assert dflt.diff(orc.compare_batch(batch, [ref], compare_cfg(CompareConfig(enable_sequences=False)))) == []


def test_speculative_dense_search_vs_oracle():
    """k_search_spec (32 queue pops in flight per dense cluster, lane-parallel exact-GT scoring) feeding the team stage: with
    AVK_DENSE_N=3 every cluster of three or more variants takes that path; AVK_NO_SPEC_SEARCH=1 keeps the team stage's own
    search covered.  Small max_branch_factor values exercise the quota-safe batch prefix."""
    p = synth.SynthParams(n_variants=2500, dense_frac=0.9, dense_mean=6.0, het_frac=0.95, phased_frac=0.2, p_repr=0.05, p_gt_err=0.05, p_fn=0.05, p_fp=0.05)
    dense_ref, dense = synth.workload_compare(50_000, p, seed=19)
    chr_ref, chr_b = synth.workload_chr20(scale=0.02, seed=31)
    for env in (dict(AVK_DENSE_N=3), dict(AVK_NO_SPEC_SEARCH=1)):
        s = _solver_with_env(**env)
        try:
            for ref, batch in ((dense_ref, dense), (chr_ref, chr_b)):
                s.set_reference([ref])
                for mbf in (50, 3, 1):
                    cfg = CompareConfig(enable_sequences=False, max_branch_factor=mbf)
                    gpu = s.compare_batch(batch, cfg)
                    cpu = orc.compare_batch(batch, [ref], compare_cfg(cfg))
                    assert gpu.diff(cpu) == [], (env, mbf)
            # sequence bundle requested: the search still runs speculatively, the team stage emits the sequences
            s.set_reference([dense_ref])
            cfg = CompareConfig(enable_sequences=True)
            off, plen = seq_offsets(dense)
            gpu, cpu = _both_compare(s, dense, [dense_ref], cfg, seq_off=off, seq_pool_len=plen)
            assert gpu.diff(cpu) == [], env
        finally:
            s.close()


def test_wfa_ed_cta_wide_paths_vs_oracle(solver):
    """Long pairs through the CTA-wide DWFA (k_wfa_ed_cta), byte-staged (the default) and with the opt-in 2-bit packed path:
    there plain ACGT pairs are packed 16 bases per word, while a single N, IUPAC code or soft-masked base anywhere in a pair
    sends it down the byte-staged path -- the reference compares raw bytes (dynamic_wfa.rs:118).  Lengths around the 16-base word and 8-diagonal group boundaries,
    one-sided empty tails, identical pairs, edit distances from 0 to a few thousand."""
    rng = np.random.default_rng(77)
    pairs = []
    for i in range(60):
        n = int(rng.integers(2500, 9000)) + (i % 17)
        a = synth.ACGT[rng.integers(0, 4, size=n)]
        kind = i % 6
        if kind == 0:
            b = a.copy()                                                       # identical
        elif kind == 1:
            b = a.copy(); idx = rng.integers(0, n, size=int(rng.integers(1, 60))); b[idx] = synth.ACGT[rng.integers(0, 4, size=idx.size)]
        elif kind == 2:
            c = int(rng.integers(0, n - 1500)); b = np.concatenate([a[:c], a[c + int(rng.integers(50, 1500)):]])          # deletion
        elif kind == 3:
            c = int(rng.integers(0, n)); b = np.concatenate([a[:c], synth.ACGT[rng.integers(0, 4, size=int(rng.integers(50, 1200)))], a[c:]])
        elif kind == 4:
            b = synth.ACGT[rng.integers(0, 4, size=int(rng.integers(2500, 4000)))]                                         # unrelated
        else:
            b = np.concatenate([a[int(rng.integers(1, 40)):], synth.ACGT[rng.integers(0, 4, size=int(rng.integers(0, 33)))]])   # shifted
        a, b = a.copy(), b.copy()
        if i % 4 == 1:
            a[int(rng.integers(0, a.size))] = ord("N")
        if i % 4 == 2:
            k = int(rng.integers(0, b.size - 20)); b[k:k + 20] = np.frombuffer(bytes(b[k:k + 20]).lower(), dtype=np.uint8)
        if i % 4 == 3 and i % 8 == 3:
            b[-1] = ord("R")
        pairs.append((a.tobytes(), b.tobytes()))
    cpu = np.array([orc.wfa_ed(a, b) for a, b in pairs], dtype=np.uint32)
    assert np.array_equal(solver.wfa_ed_batch(pairs), cpu)                    # default: byte-staged path for every pair
    assert int(cpu.max()) > 1500 and int((cpu == 0).sum()) >= 2
    s = _solver_with_env(AVK_PACKED_DWFA=1)                                   # opt-in 2-bit packed path (a device-wide switch set by avk_create)
    try:
        assert np.array_equal(s.wfa_ed_batch(pairs), cpu)
    finally:
        s.close()
        Solver(0).close()                                                     # back to the default


def _same_batch(a, b):
    for f in ("region_id", "contig", "start", "end", "var_off", "position", "variant_type", "zygosity", "raw_allele_space", "allele_off", "a0_len", "a1_len"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    n = int(a.allele_off[-1] + a.a0_len[-1] + a.a1_len[-1]) if a.n_variants else 0
    assert np.array_equal(a.allele_pool[:n], b.allele_pool[:n])


def test_build_regions_bed_multi_contig_vs_oracle(solver):
    """avk_build_regions_bed (several contigs, BED intervals, one call) against the oracle's literal restatement of the
    reference iterator: hand-made corner cases (Before / After / Overlapping variants, adjacent intervals, contig-end
    clipping), then random call sets over three contigs with random interval sets, and the built batch solved where it lies."""
    from aardvark_b200.batch import BedIntervals, CallSets
    from test_region_builder_bed import corner_case_callsets
    cs, bed, lens = corner_case_callsets()
    solver.set_reference([np.full(n, ord("A"), dtype=np.uint8) for n in lens])
    _same_batch(solver.build_regions(cs, 0, 10, first_region_id=7, bed=bed), orc.build_regions_bed(cs, lens, 10, bed, first_region_id=7))
    _same_batch(solver.build_regions(cs, 0, 10, bed=None), orc.build_regions_bed(cs, lens, 10, None))
    rng = np.random.default_rng(99)
    refs, inputs, contigs, per_contig = [], [[], []], [[], []], []
    for c, L in enumerate((60_000, 25_000, 40_000)):
        ref, (truth, query) = synth.callsets_compare(L, synth.SynthParams(n_variants=L // 150), seed=300 + c)
        refs.append(ref)
        for k, lst in enumerate((truth, query)):
            inputs[k] += lst; contigs[k] += [c] * len(lst)
        cuts = np.sort(rng.choice(np.arange(1, L), size=24, replace=False))
        ivs = [(int(cuts[2 * j]), int(cuts[2 * j + 1])) for j in range(12)]
        if c == 1:
            ivs = []                                                       # a contig without intervals is never visited
        if c == 2:
            ivs[3] = (ivs[3][0], ivs[4][0]); ivs[7] = (ivs[7][0], ivs[7][0])   # two adjacent intervals, one empty interval
        per_contig.append(ivs)
    cs, bed = CallSets(inputs, contigs=contigs), BedIntervals(per_contig)
    solver.set_reference(refs)
    for flank in (50, 0, 1000):
        gpu = solver.build_regions(cs, 0, flank, first_region_id=3, bed=bed)
        cpu = orc.build_regions_bed(cs, [r.size for r in refs], flank, bed, first_region_id=3)
        assert cpu.n_regions > 20
        _same_batch(gpu, cpu)
    # the device-built batch is resident: solve it in place and compare with the oracle on the oracle-built batch
    solver.build_regions(cs, 0, 50, bed=bed, download=False)
    cfg = CompareConfig(enable_sequences=False)
    solver.run_resident(cfg, region_metrics=True)
    cpu_b = orc.build_regions_bed(cs, [r.size for r in refs], 50, bed)
    out = CompareOutputs(cpu_b)
    solver.download(out)
    assert out.diff(orc.compare_batch(cpu_b, refs, compare_cfg(cfg))) == []


def _vcf_text(records, contig_names, sample_gt=None):
    """Call-set records [(contig, pos, a0, a1, zyg, type, raw)] -> VCF body text (one ALT per record)."""
    gt = {abi.ZYG_UNPHASED_HET: b"0/1", abi.ZYG_PHASED_HET01: b"0|1", abi.ZYG_PHASED_HET10: b"1|0", abi.ZYG_HOM_ALT: b"1/1"}
    lines = [b"##fileformat=VCFv4.2", b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE"]
    for (c, pos, a0, a1, z, t, raw) in records:
        info = b"SVTYPE=INS" if t == abi.VT_SV_INSERTION else (b"SVTYPE=DEL" if t == abi.VT_SV_DELETION else b".")
        lines.append(b"\t".join([contig_names[c].encode(), str(pos + 1).encode(), b".", a0, a1, b".", b"PASS", info, b"GT:GQ", gt[z] + b":40"]))
    return b"\n".join(lines) + b"\n"


def test_vcf_ingest_vs_oracle(solver):
    """avk_vcf_parse (one VCF record per thread: GT parsing, multi-ALT split, trimming, type inference) against the oracle's
    independent restatement: the hand-made records of tests/test_vcf_ingest.py, the error cases (same line and code), and a
    synthetic call set written out as VCF, parsed on the device, clustered by the device region builder and solved -- equal to
    the oracle on the host-built batch of the original records."""
    from aardvark_b200.batch import CallSets
    from aardvark_b200.ingest import parse_vcf_text
    from aardvark_b200.lib import AvkError
    import test_vcf_ingest as T
    names = ["chr1", "chr2"]
    for sample in (0, 1):
        for trim in (True, False):
            gpu = parse_vcf_text(solver, T.TEXT, names, sample, trim)
            cpu, err = orc.vcf_parse(T.TEXT, names, sample, trim)
            assert err == 0 and gpu.records() == cpu.records(), (sample, trim)
    for line, code in ((b"chr1\t5\t.\tA\tG\t.\t.\t.\tDP\t3\t4", 3), (b"chr1\t5\t.\tA\tG\t.\t.\t.\tGT\t0/1/1\t0/1", 4), (b"chr1\t5\t.\tA\tG\t.\t.\tSVTYPE=INV\tGT\t0/1\t0/1", 6)):
        with pytest.raises(AvkError, match=f"record 3 cannot be parsed \\(code {code}\\)"):
            parse_vcf_text(solver, T.HEADER + T.LINES[0] + b"\n" + line + b"\n", names)
    assert parse_vcf_text(solver, b"", names).n_variants == 0 and parse_vcf_text(solver, T.HEADER, names).n_variants == 0
    # synthetic call sets -> VCF text -> device parse -> device region builder -> solve
    refs, sides = [], [[], []]
    for c, L in enumerate((50_000, 30_000)):
        p = synth.SynthParams(n_variants=L // 200, sv_events=3 if c == 0 else 0, sv_min=60, sv_max=400)
        ref, (truth, query) = synth.callsets_compare(L, p, seed=500 + c)
        refs.append(ref)
        for k, lst in enumerate((truth, query)):
            sides[k] += [(c, pos, a0, a1, z, t, raw) for (pos, a0, a1, z, t, raw) in lst]
    parsed = [parse_vcf_text(solver, _vcf_text(recs, names), names, 0, True) for recs in sides]
    for recs, tab in zip(sides, parsed):
        assert tab.records() == recs                      # (the generator's records are already trimmed: raw_allele_space is kept)
    inputs = [[r[1:] for r in tab.records()] for tab in parsed]
    contigs = [[r[0] for r in tab.records()] for tab in parsed]
    solver.set_reference(refs)
    cs = CallSets(inputs, contigs=contigs)
    built = solver.build_regions(cs, 0, 50)
    host = orc.build_regions_bed(CallSets([[r[1:] for r in s] for s in sides], contigs=[[r[0] for r in s] for s in sides]), [r.size for r in refs], 50, None)
    _same_batch(built, host)
    cfg = CompareConfig(enable_sequences=False)
    assert solver.compare_batch(built, cfg).diff(orc.compare_batch(host, refs, compare_cfg(cfg))) == []


def test_bgzf_inflate_device_vs_zlib(solver):
    """avk_bgzf_inflate (one thread per BGZF member, CRC-32 checked on the device) against the original bytes that Python's
    zlib compressed: every block type and strategy of tests/test_bgzf.py, members of very different sizes in one file, the
    EOF marker, extra subfields; plain gzip, a truncated file and a flipped payload bit fail the call."""
    import zlib
    import test_bgzf as B
    from aardvark_b200.ingest import bgzf_inflate
    from aardvark_b200.lib import AvkError
    texts = B._texts()
    for name, data in sorted(texts.items()):
        for level, strategy in ((0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY)):
            gz = B.bgzf_compress(data, level, 0xff00 if level else 0xf000, strategy)
            assert bgzf_inflate(solver, gz) == data, (name, level, strategy)
    big = (texts["text"] + texts["acgt"] + texts["random"]) * 6                 # ~2.3 MB, 40 members: more than one CTA of threads
    for block in (0xff00, 4093, 257):
        assert bgzf_inflate(solver, B.bgzf_compress(big[:300000] if block == 257 else big, 6, block)) == (big[:300000] if block == 257 else big)
    gz2 = B.bgzf_member(b"hello ", extra=b"XY" + (3).to_bytes(2, "little") + b"abc") + B.bgzf_member(b"") + B.bgzf_member(b"w") + B.bgzf_member(b"orld")
    assert bgzf_inflate(solver, gz2) == b"hello world"
    assert bgzf_inflate(solver, b"") == b"" and bgzf_inflate(solver, B.BGZF_EOF) == b""
    gz = bytearray(B.bgzf_compress(texts["text"], 6))
    with pytest.raises(AvkError, match="truncated"):
        bgzf_inflate(solver, bytes(gz[:-40]))
    bad = bytearray(gz); bad[len(gz) // 2] ^= 0x55
    with pytest.raises(AvkError, match="does not inflate|BGZF member"):
        bgzf_inflate(solver, bytes(bad))
    tr = bytearray(gz); tr[len(B.bgzf_member(texts["text"][:0xff00])) - 8] ^= 1      # the first member's CRC field
    with pytest.raises(AvkError, match="code 10"):
        bgzf_inflate(solver, bytes(tr))
    assert bgzf_inflate(solver, bytes(tr), verify_crc=False) == texts["text"]
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    with pytest.raises(AvkError, match="not BGZF"):
        bgzf_inflate(solver, c.compress(texts["text"]) + c.flush())


def test_bgzf_compress_device_round_trip(solver):
    """avk_bgzf_compress: what the device writes is a BGZF file zlib reads back to the text (member layout, sizes, CRCs, EOF
    marker checked), byte for byte what the host build of the same compressor writes, and what avk_bgzf_inflate inflates again;
    labelled VCF records of a solved batch go through it and come back."""
    import test_bgzf as B
    from aardvark_b200.ingest import bgzf_inflate
    from aardvark_b200.writers import bgzf_compress, vcf_record_lines
    texts = B._texts()
    for name, data in sorted(texts.items()):
        text = data * (5 if name in ("text", "acgt") else 1)
        gz = bgzf_compress(solver, text)
        B.check_bgzf_layout(gz, text)
        assert B.gunzip_members(gz) == text, name
        assert gz == B.bgzf_compress_host(text), name
        assert bgzf_inflate(solver, gz) == text, name
    ref, batch = synth.workload_chr20(scale=0.02, seed=41)
    solver.set_reference([ref])
    out = solver.compare_batch(batch, CompareConfig(enable_sequences=False))
    lines = vcf_record_lines(batch, 1, ["chr20"], out).encode()
    assert len(lines) > 0xff00
    assert B.gunzip_members(bgzf_compress(solver, lines)) == lines


def test_vcf_parse_bgzf_vs_plain_text(solver):
    """avk_vcf_parse_bgzf = inflate + parse with the text staying on the device: the same table as avk_vcf_parse on the plain
    text (header lines skipped), for the hand-made records and for a synthetic call set of several members."""
    import test_bgzf as B
    import test_vcf_ingest as T
    from aardvark_b200.ingest import parse_vcf_bgzf, parse_vcf_text
    from aardvark_b200.lib import AvkError
    names = ["chr1", "chr2"]
    for sample, trim, block in ((0, True, 0xff00), (1, False, 97)):         # 97-byte members: records straddle member boundaries
        assert parse_vcf_bgzf(solver, B.bgzf_compress(T.TEXT, 6, block), names, sample, trim).records() == parse_vcf_text(solver, T.TEXT, names, sample, trim).records()
    recs = []
    for c, L in enumerate((400_000, 250_000)):
        ref, (truth, _q) = synth.callsets_compare(L, synth.SynthParams(n_variants=L // 150), seed=700 + c)
        recs += [(c, pos, a0, a1, z, t, raw) for (pos, a0, a1, z, t, raw) in truth]
    text = _vcf_text(recs, names)
    assert len(text) > 2 * 0xff00                                        # three members
    tab = parse_vcf_bgzf(solver, B.bgzf_compress(text, 6), names, 0, True)
    assert tab.records() == recs
    with pytest.raises(AvkError, match="record 3 cannot be parsed"):
        parse_vcf_bgzf(solver, B.bgzf_compress(T.HEADER + T.LINES[0] + b"\n" + b"chr1\t5\t.\tA\tG\t.\t.\t.\tDP\t3\t4\n"), names)
    assert parse_vcf_bgzf(solver, B.BGZF_EOF, names).n_variants == 0


def test_pipeline_bgzf_vcf_in_bgzf_outputs_out(solver):
    """The whole chain a host would run for `aardvark compare`, every stage on the device: two BGZF-compressed VCFs ->
    avk_vcf_parse_bgzf -> avk_build_regions_bed (two contigs, BED confidence regions) -> avk_compare_batch with device-side
    stratification -> summary rows and labelled VCF records -> avk_bgzf_compress.  Checked against the oracle's own chain on the
    plain text (vcf_parse -> build_regions_bed -> compare_batch): same batch, same results, same summary text; the compressed
    outputs inflate (zlib) to exactly the record lines."""
    import test_bgzf as B
    from aardvark_b200.batch import BedIntervals, CallSets, StratIntervals
    from aardvark_b200.ingest import parse_vcf_bgzf
    from aardvark_b200.writers import SummaryWriter, bgzf_compress, vcf_record_lines
    names = ["chrA", "chrB"]
    refs, sides = [], [[], []]
    for c, L in enumerate((120_000, 60_000)):
        p = synth.SynthParams(n_variants=L // 120, sv_events=2 if c == 0 else 0, sv_min=60, sv_max=300)
        ref, (truth, query) = synth.callsets_compare(L, p, seed=900 + c)
        refs.append(ref)
        for k, lst in enumerate((truth, query)):
            sides[k] += [(c, pos, a0, a1, z, t, raw) for (pos, a0, a1, z, t, raw) in lst]
    texts = [_vcf_text(recs, names) for recs in sides]
    bed = BedIntervals([[(1_000, 50_000), (52_000, 118_000)], [(0, 59_000)]])
    strat = StratIntervals([[(0, 0, 40_000), (1, 10_000, 30_000)], [(0, 30_000, 120_000)]], 2)
    # device chain
    solver.set_reference(refs)
    solver.set_stratifications(strat)
    parsed = [parse_vcf_bgzf(solver, B.bgzf_compress(t, 6, 0x4000), names, 0, True) for t in texts]
    cs = CallSets([[r[1:] for r in tab.records()] for tab in parsed], contigs=[[r[0] for r in tab.records()] for tab in parsed])
    built = solver.build_regions(cs, 0, 50, bed=bed)
    cfg = CompareConfig(enable_sequences=False)
    gpu = solver.compare_batch(built, cfg, n_strata=2, device_strata=True, containment=True)
    # oracle chain on the plain text
    o_tabs = [orc.vcf_parse(t, names, 0, True)[0] for t in texts]
    o_cs = CallSets([[r[1:] for r in tab.records()] for tab in o_tabs], contigs=[[r[0] for r in tab.records()] for tab in o_tabs])
    host = orc.build_regions_bed(o_cs, [r.size for r in refs], 50, bed)
    _same_batch(built, host)
    masks = orc.containments(host, strat)
    from aardvark_b200.batch import masks_to_membership
    so, si = masks_to_membership(masks)
    cpu = orc.compare_batch(host, refs, compare_cfg(cfg), strat_off=so, strat_idx=si, n_strata=2)
    assert gpu.diff(cpu) == []
    assert np.array_equal(gpu.containment[:host.n_regions], masks)
    sw = SummaryWriter("pipeline", strat_labels=["low", "high"])
    assert sw.summary_text(gpu.totals, gpu.strat_totals) == sw.summary_text(cpu.totals, cpu.strat_totals)
    for side in (0, 1):
        lines = vcf_record_lines(built, side, names, gpu).encode()
        assert lines == vcf_record_lines(host, side, names, cpu).encode() and lines.count(b"\n") == len(sides[side]) - _dropped(sides[side], bed)
        gz = bgzf_compress(solver, lines)
        B.check_bgzf_layout(gz, lines)
        assert B.gunzip_members(gz) == lines


def _dropped(records, bed_obj):
    """records outside every BED interval (or overlapping a boundary) never reach a region"""
    ivs = {0: [(1_000, 50_000), (52_000, 118_000)], 1: [(0, 59_000)]}
    return sum(1 for (c, pos, a0, *_r) in records if not any(s <= pos and pos + len(a0) <= e for (s, e) in ivs[c]))


def _random_adversarial_batch(n_clusters, seed, maxv=5, maxl0=3, maxins=4, wmax=160, k_inputs=2, p_copy=0.0):
    """Clusters drawn like tests/test_properties.py: overlapping records, repeated positions, ALT == REF, every zygosity,
    low-complexity windows; one window per cluster, laid end to end on one contig."""
    from aardvark_b200.types import Coordinates, PhasedZygosity, Variant, VariantType
    rng = np.random.default_rng(seed)
    zygs = [PhasedZygosity.UnphasedHeterozygous, PhasedZygosity.PhasedHet01, PhasedZygosity.PhasedHet10, PhasedZygosity.HomozygousAlternate]
    contig = bytearray()
    regions = []
    for r in range(n_clusters):
        L = int(rng.integers(60, wmax))
        alphabet = [b"ACGT", b"AC", b"A"][int(rng.integers(0, 3))]
        win = bytes(rng.choice(list(alphabet), size=L).astype(np.uint8))
        base = len(contig)
        contig += win
        sides = []
        for k_in in range(k_inputs):
            if k_in and rng.random() < p_copy:                 # merge inputs mostly agree: a copy of the first input, sometimes minus a record
                lst = list(sides[0])
                if lst and rng.random() < 0.3:
                    lst.pop(int(rng.integers(0, len(lst))))
                sides.append(lst)
                continue
            lst = []
            for p in sorted(rng.integers(10, L - 30, size=int(rng.integers(0, maxv + 1))).tolist()):
                l0 = int(rng.integers(1, maxl0 + 1)) if rng.random() < 0.4 else 1
                a0 = win[p:p + l0]
                kind = int(rng.integers(0, 4))
                if kind == 0:
                    a1 = bytes([int(rng.choice(list(b"ACGT")))]) + a0[1:]
                elif kind == 1:
                    a1 = a0[:1] + bytes(rng.choice(list(alphabet), size=int(rng.integers(1, maxins + 1))).astype(np.uint8))
                elif kind == 2:
                    a1 = a0[:1]
                else:
                    a1 = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(1, 4))).astype(np.uint8))
                vt = (VariantType.Snv if len(a0) == 1 and len(a1) == 1 else VariantType.Insertion if len(a0) == 1 else
                      VariantType.Deletion if len(a1) == 1 else VariantType.Indel)
                lst.append((Variant(0, vt, base + p, a0, a1, max(len(a0), len(a1))), zygs[int(rng.integers(0, 4))]))
            sides.append(lst)
        if not any(sides):
            continue
        if k_inputs == 2:
            regions.append(CompareRegion(len(regions), Coordinates("c", base, base + L), [v for v, _ in sides[0]], [z for _, z in sides[0]],
                                         [v for v, _ in sides[1]], [z for _, z in sides[1]]))
        else:
            from aardvark_b200.types import MultiRegion
            regions.append(MultiRegion(len(regions), Coordinates("c", base, base + L), [[v for v, _ in sd] for sd in sides], [[z for _, z in sd] for sd in sides]))
    ref = np.frombuffer(bytes(contig), dtype=np.uint8).copy()
    return ref, (RegionBatch.from_compare_regions(regions, {"c": 0}) if k_inputs == 2 else RegionBatch.from_multi_regions(regions, {"c": 0}))


@pytest.mark.parametrize("shape", [dict(seed=2718), dict(seed=31, maxv=8, maxl0=6, maxins=10, wmax=300)], ids=["small", "wide"])
def test_compare_random_adversarial_clusters_vs_oracle(solver, shape):
    """4000 randomly drawn clusters of adversarial shape through every pipeline variant: the default one (thread stage on in the
    test suite), warp kernels only, and everything with three or more variants through the speculative solver."""
    ref, batch = _random_adversarial_batch(4000, **shape)
    for mbf in (50, 2):
        cfg = CompareConfig(enable_sequences=False, max_branch_factor=mbf)
        cpu = orc.compare_batch(batch, [ref], compare_cfg(cfg))
        assert int((cpu.status[:batch.n_regions] == 0).sum()) > 0.9 * batch.n_regions
        solver.set_reference([ref])
        assert solver.compare_batch(batch, cfg).diff(cpu) == [], ("default", mbf)
        for env in (dict(AVK_NO_THREAD_STAGE=1), dict(AVK_NO_THREAD_STAGE=1, AVK_DENSE_N=3), dict(AVK_DENSE_N=3, AVK_THREAD_POP_BUDGET=6)):
            s = _solver_with_env(**env)
            try:
                s.set_reference([ref])
                assert s.compare_batch(batch, cfg).diff(cpu) == [], (env, mbf)
            finally:
                s.close()


def test_compare_random_adversarial_clusters_exact_shortcut_and_sequences_vs_oracle(solver):
    """The configurations that bypass the fast kernels (enable_exact_shortcut: waffle_solver.rs:171-199; enable_sequences: the
    sequence bundle) on adversarial clusters whose query side is mostly a copy of the truth side, so that the shortcut fires --
    including copies that contain ALT == REF records, repeated positions and overlapping records."""
    ref, batch = _random_adversarial_batch(2500, seed=99, k_inputs=2, p_copy=0.6)
    solver.set_reference([ref])
    off, plen = seq_offsets(batch)
    for mbf in (50, 2):
        for shortcut in (True, False):
            cfg = CompareConfig(enable_sequences=False, enable_exact_shortcut=shortcut, max_branch_factor=mbf)
            cpu = orc.compare_batch(batch, [ref], compare_cfg(cfg))
            assert solver.compare_batch(batch, cfg).diff(cpu) == [], (mbf, shortcut)
        cs = CompareConfig(enable_sequences=True, enable_exact_shortcut=True, max_branch_factor=mbf)
        cpu = orc.compare_batch(batch, [ref], compare_cfg(cs), seq_off=off, seq_pool_len=plen)
        assert solver.compare_batch(batch, cs, seq_off=off, seq_pool_len=plen).diff(cpu) == [], (mbf, "sequences")


def test_merge_random_adversarial_clusters_vs_oracle(solver):
    """solve_merge_region on randomly drawn clusters of four inputs that mostly agree (copies of the first input, sometimes
    minus a record) or are drawn independently -- overlapping records, repeated positions, ALT == REF -- under every merge
    configuration."""
    ref, batch = _random_adversarial_batch(2500, seed=77, k_inputs=4, p_copy=0.7)
    solver.set_reference([ref])
    for cfg in (MergeConfig(majority_voting_enabled=True), MergeConfig(no_conflict_enabled=True), MergeConfig(conflict_selection=1), MergeConfig()):
        gpu = solver.merge_batch(batch, cfg)
        cpu = orc.merge_batch(batch, [ref], merge_cfg(cfg))
        assert gpu.diff(cpu) == [], cfg
    assert len(set(cpu.classification[:batch.n_regions].tolist())) >= 2
