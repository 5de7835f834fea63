// aardvark_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain C++17 restatement of the aardvark v0.10.5 hot path (haplotype-level
// comparison of truth vs query variant clusters).  It exists so that the CUDA
// product in aardvark_b200/csrc can be checked bit for bit, and so that
// bench.py has a CPU baseline ("port").  Nothing in the product path may call
// into this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs do.
//
// PARITY PINNING: the reference cannot be built here (no cargo/rustc), so this
// restatement is pinned against every golden vector of the reference's own
// unit tests and doctests for the path (tests/test_oracle_golden.py lists them
// with file:line).  FASTA parsing of the un-vendored rust-lib-reference-genome
// crate is NOT on this path (contigs arrive as raw byte arrays) -- that part
// is "parity unpinned" and out of scope.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference).  The code is written from the behaviour, not translated
// line by line: sequences are std::vector<uint8_t>, the priority queue is a
// binary heap keyed by the reference's (unique) priority tuples.

#include "oracle.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <queue>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

using Seq = std::vector<uint8_t>;

struct Work {
    uint64_t alignments = 0, cells = 0, matched = 0, search_pops = 0, exact_pops = 0, alg_bytes = 0;
    void add(const Work &o) {
        alignments += o.alignments; cells += o.cells; matched += o.matched;
        search_pops += o.search_pops; exact_pops += o.exact_pops; alg_bytes += o.alg_bytes;
    }
};
static thread_local Work *tl_work = nullptr;

// ---------------------------------------------------------------------------
// DWFALite -- src/dwfa/dynamic_wfa.rs:22-276
// ---------------------------------------------------------------------------
enum class DErr { Ok, MaxEditDistance, AlreadyFinalized };

struct DWFALite {
    size_t ed = 0;                    // :25
    std::vector<size_t> wf{0};        // :34, starts as [0] (:45)
    bool finalized = false;           // :36
    size_t max_ed = SIZE_MAX;         // :38,47

    // extend(): :94-130.  wf[i] = bases consumed in `other`; baseline offset = d + ed - i (:114)
    void extend(const uint8_t *b, size_t lb, const uint8_t *o, size_t lo) {
        const size_t n = wf.size();
        size_t matched = 0;
        for (size_t i = 0; i < n; ++i) {
            size_t d = wf[i];
            size_t boff = d + ed - i;
            while (boff < lb && d < lo && b[boff] == o[d]) { ++d; ++boff; }
            matched += d - wf[i];
            wf[i] = d;
        }
        if (tl_work) { tl_work->cells += n; tl_work->matched += matched; }
    }

    // increase_edit_distance(): :140-173.  NOTE ed is incremented before the max check (:146-149).
    DErr increase(const uint8_t *b, size_t lb, const uint8_t *o, size_t lo) {
        if (finalized) return DErr::AlreadyFinalized;
        ed += 1;
        if (ed > max_ed) return DErr::MaxEditDistance;
        std::vector<size_t> nw(wf.size() + 2, 0);
        for (size_t i = 0; i < wf.size(); ++i) {
            size_t d = wf[i];
            nw[i] = std::max(nw[i], d);             // :158
            nw[i + 1] = std::max(nw[i + 1], d + 1); // :161
            nw[i + 2] = std::max(nw[i + 2], d + 1); // :164
        }
        wf.swap(nw);
        extend(b, lb, o, lo);
        return DErr::Ok;
    }

    size_t max_baseline() const {   // :201-208
        size_t m = 0;
        for (size_t i = 0; i < wf.size(); ++i) m = std::max(m, wf[i] + ed - i);
        return m;
    }
    size_t max_other() const {      // :212-215
        size_t m = 0;
        for (size_t v : wf) m = std::max(m, v);
        return m;
    }
    bool reached_full(size_t lb, size_t lo) const {  // :237-245
        for (size_t i = 0; i < wf.size(); ++i)
            if (wf[i] + ed - i >= lb && wf[i] >= lo) return true;
        return false;
    }

    // update(): :68-84 -- stop when a diagonal touches the end of EITHER sequence (:77)
    DErr update(const uint8_t *b, size_t lb, const uint8_t *o, size_t lo) {
        if (finalized) return DErr::AlreadyFinalized;
        extend(b, lb, o, lo);
        while (!(max_baseline() >= lb) && !(max_other() >= lo)) {
            DErr e = increase(b, lb, o, lo);
            if (e != DErr::Ok) return e;
        }
        return DErr::Ok;
    }

    // finalize(): :183-198 -- stop when one diagonal is at the end of BOTH sequences (:192)
    DErr finalize(const uint8_t *b, size_t lb, const uint8_t *o, size_t lo) {
        if (finalized) return DErr::AlreadyFinalized;
        if (tl_work) { tl_work->alignments += 1; tl_work->alg_bytes += (lb + 3) / 4 + (lo + 3) / 4 + 8; }   // 2-bit bases + 8
        extend(b, lb, o, lo);
        while (!reached_full(lb, lo)) {
            DErr e = increase(b, lb, o, lo);
            if (e != DErr::Ok) return e;
        }
        finalized = true;
        return DErr::Ok;
    }
};

// wfa_ed(): src/util/sequence_alignment.rs:9-13
static size_t wfa_ed(const uint8_t *a, size_t la, const uint8_t *b, size_t lb) {
    DWFALite d;
    d.finalize(a, la, b, lb);
    return d.ed;
}

// edit_distance(): src/util/sequence_alignment.rs:20-51 (two-row DP; same value as wfa_ed)
static size_t edit_distance(const uint8_t *v1, size_t l1, const uint8_t *v2, size_t l2) {
    std::vector<size_t> row(l1 + 1, 0), prev(l1 + 1);
    for (size_t j = 0; j <= l1; ++j) prev[j] = j;
    for (size_t i = 0; i < l2; ++i) {
        row[0] = i + 1;
        for (size_t j = 0; j < l1; ++j) {
            size_t a = prev[j + 1] + 1, c = row[j] + 1, dg = prev[j] + (v1[j] == v2[i] ? 0 : 1);
            row[j + 1] = std::min(a, std::min(c, dg));
        }
        row.swap(prev);
    }
    return prev[l1];
}

// ---------------------------------------------------------------------------
// Variant view -- the fields of src/data_types/variants.rs:73-91 read on the path
// ---------------------------------------------------------------------------
struct Var {
    uint32_t pos;
    uint8_t type;
    const uint8_t *a0; uint32_t l0;
    const uint8_t *a1; uint32_t l1;
    uint32_t raw;
    size_t ref_len() const { return l0; }                                 // variants.rs:437-439
    size_t alt_ed() const { return wfa_ed(a0, l0, a1, l1); }              // variants.rs:413-415
};

enum : uint8_t { AL_UNKNOWN = 0, AL_REF = 1, AL_ALT = 2 };   // phase_enums.rs:9-16
static inline uint8_t al_count(uint8_t a) { return a == AL_ALT ? 1 : 0; }   // :20-26

static inline bool zyg_is_het(uint8_t z) {        // phase_enums.rs:63-73
    return z == AVK_ZYG_UNPHASED_HET || z == AVK_ZYG_PHASED_HET01 || z == AVK_ZYG_PHASED_HET10;
}
static inline void zyg_decompose(uint8_t z, uint8_t &a1, uint8_t &a2) {   // phase_enums.rs:91-100
    switch (z) {
        case AVK_ZYG_HOM_REF: a1 = AL_REF; a2 = AL_REF; break;
        case AVK_ZYG_UNPHASED_HET:
        case AVK_ZYG_PHASED_HET01: a1 = AL_REF; a2 = AL_ALT; break;
        case AVK_ZYG_PHASED_HET10: a1 = AL_ALT; a2 = AL_REF; break;
        case AVK_ZYG_HOM_ALT: a1 = AL_ALT; a2 = AL_ALT; break;
        default: a1 = AL_UNKNOWN; a2 = AL_UNKNOWN; break;
    }
}
static inline uint8_t zyg_count(uint8_t z) {      // phase_enums.rs:103-112
    if (zyg_is_het(z)) return 1;
    if (z == AVK_ZYG_HOM_ALT) return 2;
    return 0;
}

// ---------------------------------------------------------------------------
// HaplotypeTracker / HaplotypeDWFA -- src/dwfa/haplotype_dwfa.rs
// ---------------------------------------------------------------------------
struct Tracker {                       // :145-154
    size_t ref_pos = 0;
    std::vector<uint8_t> alleles;
    Seq seq;
    size_t skip = 0;
    explicit Tracker(size_t start = 0) : ref_pos(start) {}

    void copy_reference(const uint8_t *ref, size_t end) {    // :218-227
        if (ref_pos < end) {
            seq.insert(seq.end(), ref + ref_pos, ref + end);
            ref_pos = end;
        }
    }
    // extend_variant(): :175-212.  returns 0/1 success, -1 for Allele::Unknown
    int extend_variant(const uint8_t *ref, const Var &v, uint8_t allele, bool has_ext, size_t ext) {
        size_t vstart = v.pos;
        copy_reference(ref, vstart);
        int success;
        if (allele == AL_UNKNOWN) return -1;
        if (allele == AL_REF) {
            success = 1;
        } else {
            if (ref_pos <= vstart) {                 // compatible (:189)
                seq.insert(seq.end(), v.a1, v.a1 + v.l1);
                ref_pos = vstart + v.ref_len();
                success = 1;
            } else {                                 // incompatible: cost = edit_distance(a0,a1) (:199)
                skip += edit_distance(v.a0, v.l0, v.a1, v.l1);
                success = 0;
            }
        }
        alleles.push_back(allele);                   // :204
        if (has_ext) copy_reference(ref, ext);       // :207-209
        return success;
    }
};

struct HapDWFA {                      // :17-24
    Tracker truth, query;
    DWFALite dwfa;
    HapDWFA(size_t start, size_t max_ed) : truth(start), query(start) { dwfa.max_ed = max_ed; }

    // extend_variant(): :46-67.  *derr receives a DWFA error (success is then meaningless)
    int extend_variant(const uint8_t *ref, bool is_truth, const Var &v, uint8_t allele,
                       bool has_sync, size_t sync, DErr *derr) {
        int success;
        if (is_truth) {
            if (has_sync) query.copy_reference(ref, sync);
            success = truth.extend_variant(ref, v, allele, has_sync, sync);
        } else {
            if (has_sync) truth.copy_reference(ref, sync);
            success = query.extend_variant(ref, v, allele, has_sync, sync);
        }
        if (success < 0) { *derr = DErr::Ok; return -1; }
        *derr = dwfa.update(truth.seq.data(), truth.seq.size(), query.seq.data(), query.seq.size());  // :74
        return success;
    }
    DErr finalize_dwfa(const uint8_t *ref, size_t region_end) {   // :84-95
        truth.copy_reference(ref, region_end);
        query.copy_reference(ref, region_end);
        DErr e = dwfa.update(truth.seq.data(), truth.seq.size(), query.seq.data(), query.seq.size());
        if (e != DErr::Ok) return e;
        return dwfa.finalize(truth.seq.data(), truth.seq.size(), query.seq.data(), query.seq.size());
    }
    bool is_synchronized() const {     // :99-112
        return dwfa.ed == 0 && truth.seq.size() == query.seq.size() && truth.ref_pos == query.ref_pos;
    }
    size_t set_alleles() const { return truth.alleles.size() + query.alleles.size(); }   // :120-122
    size_t total_cost() const { return dwfa.ed + truth.skip + query.skip; }               // :130-132
};

// order_variants(): src/query_optimizer.rs:372-381 (stable sort => truth before query on ties)
struct OrderEnt { uint32_t idx; bool is_truth; };
static std::vector<OrderEnt> order_variants(const std::vector<Var> &t, const std::vector<Var> &q) {
    std::vector<OrderEnt> r;
    r.reserve(t.size() + q.size());
    for (uint32_t i = 0; i < t.size(); ++i) r.push_back({i, true});
    for (uint32_t i = 0; i < q.size(); ++i) r.push_back({i, false});
    std::stable_sort(r.begin(), r.end(), [&](const OrderEnt &a, const OrderEnt &b) {
        uint32_t pa = a.is_truth ? t[a.idx].pos : q[a.idx].pos;
        uint32_t pb = b.is_truth ? t[b.idx].pos : q[b.idx].pos;
        return pa < pb;
    });
    return r;
}

// ---------------------------------------------------------------------------
// optimize_sequences -- src/query_optimizer.rs:166-365
// ---------------------------------------------------------------------------
struct OptHaps {                      // OptimizedHaplotypes :67-92
    std::vector<uint8_t> truth_zyg, query_zyg;
    Seq truth_seq1, truth_seq2, query_seq1, query_seq2;
    size_t ed1 = 0, ed2 = 0, truth_vs1 = 0, truth_vs2 = 0, query_vs1 = 0, query_vs2 = 0;
    bool is_exact_match() const { return ed1 + ed2 + truth_vs1 + truth_vs2 + query_vs1 + query_vs2 == 0; }  // :96-98
};

struct CmpNode {                      // ComparisonNode :406-413
    uint64_t id;
    HapDWFA h1, h2;
    CmpNode(uint64_t id_, size_t start) : id(id_), h1(start, SIZE_MAX), h2(start, SIZE_MAX) {}
    size_t total_cost() const { return h1.total_cost() + h2.total_cost(); }   // :465-467
    size_t set_alleles() const { return h1.set_alleles(); }                    // :478-481
};

static bool convert_alleles_to_zyg(const std::vector<uint8_t> &a1, const std::vector<uint8_t> &a2,
                                   std::vector<uint8_t> &out) {   // :388-402
    out.clear();
    for (size_t i = 0; i < a1.size(); ++i) {
        if (a1[i] == AL_REF && a2[i] == AL_ALT) out.push_back(AVK_ZYG_PHASED_HET01);
        else if (a1[i] == AL_ALT && a2[i] == AL_REF) out.push_back(AVK_ZYG_PHASED_HET10);
        else if (a1[i] == AL_ALT && a2[i] == AL_ALT) out.push_back(AVK_ZYG_HOM_ALT);
        else return false;   // panic!("no impl")
    }
    return true;
}

// returns AVK_ST_*; when stop_on_nonzero the search returns as soon as it is known
// whether the minimum cost is zero (merge only needs is_exact_match of result[0]).
static int optimize_sequences(const uint8_t *ref, size_t start, size_t end,
                              const std::vector<Var> &tv, const std::vector<uint8_t> &tz,
                              const std::vector<Var> &qv, const std::vector<uint8_t> &qz,
                              size_t max_branch_factor, std::vector<OptHaps> &out) {
    out.clear();
    if (max_branch_factor == 0) return AVK_ST_BAD_INPUT;           // :177
    std::vector<OrderEnt> order = order_variants(tv, qv);
    const size_t N = order.size();

    struct QEnt { size_t cost; uint64_t id; size_t slot; };
    auto worse = [](const QEnt &a, const QEnt &b) {            // min (cost, id) first  :417,470-475
        if (a.cost != b.cost) return a.cost > b.cost;
        return a.id > b.id;
    };
    std::priority_queue<QEnt, std::vector<QEnt>, decltype(worse)> pq(worse);
    std::vector<std::unique_ptr<CmpNode>> pool;
    uint64_t next_id = 0;
    pool.emplace_back(new CmpNode(next_id++, start));
    pq.push({pool[0]->total_cost(), pool[0]->id, 0});

    size_t best = SIZE_MAX;
    std::vector<std::unique_ptr<CmpNode>> best_nodes;
    std::vector<size_t> bucket(N + 1, 0);                         // :200

    while (!pq.empty()) {
        QEnt top = pq.top(); pq.pop();
        std::unique_ptr<CmpNode> cur = std::move(pool[top.slot]);
        if (tl_work) tl_work->search_pops += 1;
        if (cur->total_cost() > best) continue;                   // :204 (strict)
        size_t oi = cur->set_alleles();                           // :210
        if (bucket[oi] >= max_branch_factor) continue;            // :222
        bucket[oi] += 1;
        if (oi == N) {                                            // :227-247
            DErr e1 = cur->h1.finalize_dwfa(ref, end);
            if (e1 != DErr::Ok) return AVK_ST_NO_RESULT;
            DErr e2 = cur->h2.finalize_dwfa(ref, end);
            if (e2 != DErr::Ok) return AVK_ST_NO_RESULT;
            size_t c = cur->total_cost();
            if (c < best) { best = c; best_nodes.clear(); best_nodes.push_back(std::move(cur)); }
            else if (c == best) best_nodes.push_back(std::move(cur));
            continue;
        }
        const OrderEnt oe = order[oi];
        const Var &v = oe.is_truth ? tv[oe.idx] : qv[oe.idx];
        const uint8_t z = oe.is_truth ? tz[oe.idx] : qz[oe.idx];
        size_t sync;                                              // :258-265
        if (oi == N - 1) sync = end;
        else { const OrderEnt &ne = order[oi + 1]; sync = ne.is_truth ? tv[ne.idx].pos : qv[ne.idx].pos; }

        auto extend_both = [&](CmpNode &n, uint8_t a1, uint8_t a2) -> bool {   // :443-451
            DErr de;
            int s1 = n.h1.extend_variant(ref, oe.is_truth, v, a1, true, sync, &de);
            if (s1 < 0 || de != DErr::Ok) return false;
            int s2 = n.h2.extend_variant(ref, oe.is_truth, v, a2, true, sync, &de);
            if (s2 < 0 || de != DErr::Ok) return false;
            return true;
        };
        auto push = [&](std::unique_ptr<CmpNode> n) {
            QEnt e{n->total_cost(), n->id, pool.size()};
            pool.push_back(std::move(n));
            pq.push(e);
        };

        if (zyg_is_het(z)) {
            if (!oe.is_truth || z == AVK_ZYG_UNPHASED_HET) {      // :269-293: both orientations, new ids
                const uint8_t ext[2][2] = {{AL_REF, AL_ALT}, {AL_ALT, AL_REF}};
                for (int k = 0; k < 2; ++k) {
                    std::unique_ptr<CmpNode> nn(new CmpNode(*cur));
                    nn->id = next_id++;
                    if (!extend_both(*nn, ext[k][0], ext[k][1])) return AVK_ST_NO_RESULT;
                    push(std::move(nn));
                }
            } else {                                               // :294-312 phased truth het, id kept
                uint8_t a1 = (z == AVK_ZYG_PHASED_HET01) ? AL_REF : AL_ALT;
                uint8_t a2 = (z == AVK_ZYG_PHASED_HET01) ? AL_ALT : AL_REF;
                if (!extend_both(*cur, a1, a2)) return AVK_ST_NO_RESULT;
                push(std::move(cur));
            }
        } else {
            if (z != AVK_ZYG_HOM_ALT) return AVK_ST_BAD_ZYGOSITY;   // assert_eq! :315 (process panic in the reference)
            if (!extend_both(*cur, AL_ALT, AL_ALT)) return AVK_ST_NO_RESULT;
            push(std::move(cur));
        }
    }
    if (best_nodes.empty()) return AVK_ST_NO_RESULT;              // :331

    for (auto &bn : best_nodes) {                                  // :334-363
        OptHaps o;
        if (!convert_alleles_to_zyg(bn->h1.truth.alleles, bn->h2.truth.alleles, o.truth_zyg)) return AVK_ST_BAD_ZYGOSITY;
        if (!convert_alleles_to_zyg(bn->h1.query.alleles, bn->h2.query.alleles, o.query_zyg)) return AVK_ST_BAD_ZYGOSITY;
        o.truth_seq1 = bn->h1.truth.seq; o.truth_seq2 = bn->h2.truth.seq;
        o.query_seq1 = bn->h1.query.seq; o.query_seq2 = bn->h2.query.seq;
        o.ed1 = bn->h1.dwfa.ed; o.ed2 = bn->h2.dwfa.ed;
        o.truth_vs1 = bn->h1.truth.skip; o.truth_vs2 = bn->h2.truth.skip;
        o.query_vs1 = bn->h1.query.skip; o.query_vs2 = bn->h2.query.skip;
        out.push_back(std::move(o));
    }
    return AVK_ST_OK;
}

// ---------------------------------------------------------------------------
// optimize_gt_alleles -- src/exact_gt_optimizer.rs:108-357
// ---------------------------------------------------------------------------
struct OptAlleles { std::vector<uint8_t> truth_alleles, query_alleles; size_t num_errors = 0; };

struct ExactNode {                    // ExactMatchNode :361-368, DWFA max ED 0 (:380)
    uint64_t id;
    HapDWFA h;
    size_t errors = 0;
    ExactNode(uint64_t id_, size_t start) : id(id_), h(start, 0) {}
    bool is_exact() const { return h.dwfa.ed == 0; }           // :442-444
    // extend_variant(): :395-414.  MaxEditDistance is tolerated (:399-409, :482-488)
    // returns 1/0 for Ok(bool), -1 for a hard error
    int extend_variant(const uint8_t *ref, bool is_truth, const Var &v, uint8_t allele, size_t sync, bool is_error) {
        DErr de;
        int s = h.extend_variant(ref, is_truth, v, allele, true, sync, &de);
        if (s < 0) return -1;
        int ext = s;
        if (de == DErr::MaxEditDistance) ext = 0;
        else if (de != DErr::Ok) return -1;
        if (is_error) errors += 1;
        return ext;
    }
};

static int optimize_gt_alleles(const uint8_t *ref, size_t start, size_t end,
                               const std::vector<Var> &tv, const std::vector<uint8_t> &ta,
                               const std::vector<Var> &qv, const std::vector<uint8_t> &qa,
                               OptAlleles &out, uint64_t max_expansions = AVK_EXACT_GT_DEFAULT_MAX_EXPANSIONS) {
    uint64_t expansions = 0;
    std::vector<OrderEnt> order = order_variants(tv, qv);
    const size_t N = order.size();

    struct QEnt { size_t errors; size_t good; uint64_t id; size_t slot; };
    // priority (Reverse(errors), set - errors, Reverse(id)) max-first :372,452-458
    auto worse = [](const QEnt &a, const QEnt &b) {
        if (a.errors != b.errors) return a.errors > b.errors;
        if (a.good != b.good) return a.good < b.good;
        return a.id > b.id;
    };
    std::vector<QEnt> heap;    // kept as a heap; rebuilt on the auto-fail filter
    std::vector<std::unique_ptr<ExactNode>> pool;
    uint64_t next_id = 0;
    pool.emplace_back(new ExactNode(next_id++, start));
    heap.push_back({0, 0, 0, 0});

    size_t best_err = SIZE_MAX;
    std::unique_ptr<ExactNode> best_node;
    size_t min_allele_sync = 0;
    const size_t auto_fail_threshold = 500;      // :160
    size_t auto_fail_index = 0, auto_fail_counts = 0;

    auto push = [&](std::unique_ptr<ExactNode> n) {
        QEnt e{n->errors, n->h.set_alleles() - n->errors, n->id, pool.size()};
        pool.push_back(std::move(n));
        heap.push_back(e);
        std::push_heap(heap.begin(), heap.end(), worse);
    };

    while (!heap.empty()) {
        std::pop_heap(heap.begin(), heap.end(), worse);
        QEnt top = heap.back(); heap.pop_back();
        std::unique_ptr<ExactNode> cur = std::move(pool[top.slot]);
        if (tl_work) tl_work->exact_pops += 1;
        if (cur->errors >= best_err) continue;                    // :169 (non-strict)
        // the 300 s wall-clock bail (:174-176) is non-deterministic; its stand-in is a cap on the nodes expanded by one call,
        // tested where the reference reads its clock
        if (++expansions > max_expansions) return AVK_ST_TIMEOUT;
        size_t oi = cur->h.set_alleles();
        if (oi == N) {                                            // :180-192
            DErr e = cur->h.finalize_dwfa(ref, end);
            if (e != DErr::Ok && e != DErr::MaxEditDistance) return AVK_ST_NO_RESULT;
            if (cur->is_exact() && cur->errors < best_err) { best_err = cur->errors; best_node = std::move(cur); }
            continue;
        }
        if (oi < min_allele_sync) continue;                       // :194-197
        if (cur->h.is_synchronized()) {                           // :206-217
            min_allele_sync = oi;
            auto_fail_counts = 0;
            auto_fail_index = min_allele_sync;
        }
        const OrderEnt oe = order[oi];
        const Var &v = oe.is_truth ? tv[oe.idx] : qv[oe.idx];
        const uint8_t al = oe.is_truth ? ta[oe.idx] : qa[oe.idx];
        size_t sync;
        if (oi == N - 1) sync = end;
        else { const OrderEnt &ne = order[oi + 1]; sync = ne.is_truth ? tv[ne.idx].pos : qv[ne.idx].pos; }

        if (al == AL_UNKNOWN) return AVK_ST_BAD_ZYGOSITY;         // :256
        if (al == AL_REF) {                                       // :257-273 move, id kept
            int s = cur->extend_variant(ref, oe.is_truth, v, AL_REF, sync, false);
            if (s < 0) return AVK_ST_NO_RESULT;
            if (s == 1 && cur->is_exact()) push(std::move(cur));
        } else {                                                  // :274-306
            const uint8_t ext_al[2] = {AL_REF, AL_ALT};
            const bool ext_err[2] = {true, false};
            for (int k = 0; k < 2; ++k) {
                if (oi < auto_fail_index && ext_al[k] != AL_REF) continue;   // :282-285
                std::unique_ptr<ExactNode> nn(new ExactNode(*cur));
                nn->id = next_id++;
                int s = nn->extend_variant(ref, oe.is_truth, v, ext_al[k], sync, ext_err[k]);
                if (s < 0) return AVK_ST_NO_RESULT;
                if (s == 1 && nn->is_exact()) push(std::move(nn));
            }
        }
        auto_fail_counts += 1;                                    // :310-339
        if (auto_fail_counts >= auto_fail_threshold) {
            if (auto_fail_index >= N) return AVK_ST_NO_RESULT;    // reference would index out of bounds (panic)
            const OrderEnt af = order[auto_fail_index];
            std::vector<QEnt> kept;
            for (const QEnt &e : heap) {
                const ExactNode &n = *pool[e.slot];
                const std::vector<uint8_t> &als = af.is_truth ? n.h.truth.alleles : n.h.query.alleles;
                uint8_t a = af.idx < als.size() ? als[af.idx] : (uint8_t)AL_REF;
                if (a == AL_REF) kept.push_back(e);
                else pool[e.slot].reset();
            }
            heap.swap(kept);
            std::make_heap(heap.begin(), heap.end(), worse);
            auto_fail_index += 1;
            auto_fail_counts = 0;
        }
    }
    if (!best_node) return AVK_ST_NO_RESULT;                      // :345-348
    out.truth_alleles = best_node->h.truth.alleles;
    out.query_alleles = best_node->h.query.alleles;
    out.num_errors = best_node->errors;
    return AVK_ST_OK;
}

// ---------------------------------------------------------------------------
// Metrics -- src/data_types/{grouped_metrics,summary_metrics,variant_metrics}.rs
// ---------------------------------------------------------------------------
struct GroupTypeMetrics {             // grouped_metrics.rs:31-37 flattened
    uint64_t m[AVK_N_GROUPS][AVK_N_METRICS];
    uint16_t mask = 0;                // which variant types have an entry
    GroupTypeMetrics() { std::memset(m, 0, sizeof(m)); }
    uint64_t *group(int vt /* -1 joint */) {
        if (vt < 0) return m[0];
        mask |= (uint16_t)(1u << vt);
        return m[1 + vt];
    }
};

// GroupMetrics::add_truth_zygosity  grouped_metrics.rs:183-227.  returns status
static int gm_add_truth_zygosity(uint64_t *g, uint64_t w, uint8_t exp, uint8_t obs) {
    if (exp == 0) return AVK_ST_TRUTH_FP;     // ensure!(expected > 0)
    if (exp < obs) return AVK_ST_TRUTH_FP;    // bail!("No implementation for truth false positives")
    if (exp == obs) {
        g[AVK_M_HAP + 0] += exp;
        g[AVK_M_WEIGHTED_HAP + 0] += (uint64_t)exp * w;
        g[AVK_M_GT + 0] += 1;
    } else {
        g[AVK_M_HAP + 0] += obs;
        g[AVK_M_HAP + 1] += (uint64_t)(exp - obs);
        g[AVK_M_WEIGHTED_HAP + 0] += (uint64_t)obs * w;
        g[AVK_M_WEIGHTED_HAP + 1] += (uint64_t)(exp - obs) * w;
        g[AVK_M_GT + 1] += 1;
        if (obs > 0) g[AVK_M_GT_TRUTH_FN_GT] += 1;
    }
    return AVK_ST_OK;
}
// GroupMetrics::add_query_zygosity  grouped_metrics.rs:234-249 (exact shortcut only)
static int gm_add_query_zygosity(uint64_t *g, uint64_t w, uint8_t exp, uint8_t obs) {
    if (exp == 0 || exp != obs) return AVK_ST_TRUTH_FP;
    g[AVK_M_HAP + 2] += exp;
    g[AVK_M_WEIGHTED_HAP + 2] += (uint64_t)exp * w;
    g[AVK_M_GT + 2] += 1;
    return AVK_ST_OK;
}
// GroupMetrics::add_swap_benchmark  grouped_metrics.rs:268-277, summary_metrics.rs:35-38,111-114
static void gm_add_swap(uint64_t *g, const uint64_t *o) {
    g[AVK_M_GT + 2] = o[AVK_M_GT + 0];  g[AVK_M_GT + 3] = o[AVK_M_GT + 1];
    g[AVK_M_GT_QUERY_FP_GT] = o[AVK_M_GT_TRUTH_FN_GT];
    g[AVK_M_HAP + 2] = o[AVK_M_HAP + 0];  g[AVK_M_HAP + 3] = o[AVK_M_HAP + 1];
    g[AVK_M_WEIGHTED_HAP + 2] = o[AVK_M_WEIGHTED_HAP + 0];  g[AVK_M_WEIGHTED_HAP + 3] = o[AVK_M_WEIGHTED_HAP + 1];
}
static void add4(uint64_t *g, int base, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    g[base] += a; g[base + 1] += b; g[base + 2] += c; g[base + 3] += d;
}

// VariantMetrics::new classification  variant_metrics.rs:43-71
static int classify(uint8_t exp, uint8_t obs, uint8_t &cls) {
    if (exp > 2 || obs > 2) return AVK_ST_TRUTH_FP;
    if (exp < obs) cls = AVK_CLASS_FP;
    else if (exp == obs) { if (exp == 0) return AVK_ST_TRUTH_FP; cls = AVK_CLASS_TP; }
    else cls = AVK_CLASS_FN;
    return AVK_ST_OK;
}
static uint8_t toggle_class(uint8_t c) {          // variant_metrics.rs:77-101
    if (c == AVK_CLASS_FN) return AVK_CLASS_FP;
    if (c == AVK_CLASS_FP) return AVK_CLASS_FN;
    return c;
}

struct VarMetric { uint8_t expected, observed, cls; };

struct Benchmark {                    // CompareBenchmark compare_benchmark.rs:9-33
    size_t ed1 = 0, ed2 = 0;
    GroupTypeMetrics gm;
    std::vector<VarMetric> truth_data, query_data;
    bool has_seqs = false;
    Seq ref_seq, truth_seq1, truth_seq2, query_seq1, query_seq2;

    int add_truth_zygosity(const Var &v, uint8_t exp, uint8_t obs) {     // :61-70 + grouped_metrics.rs:45-61
        if (exp == 0) return AVK_ST_TRUTH_FP;
        uint64_t w = v.alt_ed();
        int st = gm_add_truth_zygosity(gm.group(-1), w, exp, obs);
        if (st) return st;
        st = gm_add_truth_zygosity(gm.group(v.type), w, exp, obs);
        if (st) return st;
        VarMetric vm{exp, obs, 0};
        st = classify(exp, obs, vm.cls);
        if (st) return st;
        truth_data.push_back(vm);
        return AVK_ST_OK;
    }
    int add_query_zygosity(const Var &v, uint8_t exp, uint8_t obs) {     // :77-88 + grouped_metrics.rs:68-81
        uint64_t w = v.alt_ed();
        int st = gm_add_query_zygosity(gm.group(-1), w, exp, obs);
        if (st) return st;
        st = gm_add_query_zygosity(gm.group(v.type), w, exp, obs);
        if (st) return st;
        uint8_t c;
        st = classify(exp, obs, c);
        if (st) return st;
        query_data.push_back({obs, exp, toggle_class(c)});
        return AVK_ST_OK;
    }
    void add_swap(const Benchmark &o) {                                  // :109-123 + grouped_metrics.rs:113-120
        gm_add_swap(gm.m[0], o.gm.m[0]);
        for (int t = 0; t < AVK_N_VARIANT_TYPES; ++t)
            if (o.gm.mask & (1u << t)) gm_add_swap(gm.group(t), o.gm.m[1 + t]);
        for (const VarMetric &vm : o.truth_data)
            query_data.push_back({vm.observed, vm.expected, toggle_class(vm.cls)});
    }
};

// ---------------------------------------------------------------------------
// waffle solver -- src/waffle_solver.rs
// ---------------------------------------------------------------------------
static const int SUPPORTED_TYPES[8] = {    // :82-91 (this ORDER is the iteration order)
    AVK_VT_SNV, AVK_VT_INSERTION, AVK_VT_DELETION, AVK_VT_INDEL,
    AVK_VT_TR_CONTRACTION, AVK_VT_TR_EXPANSION, AVK_VT_SV_DELETION, AVK_VT_SV_INSERTION};

// generate_allele_sequence(): :726-778
static int generate_allele_sequence(const uint8_t *ref, size_t start, size_t end,
                                    const std::vector<const Var *> &vars, const std::vector<uint8_t> &alleles,
                                    Seq &seq, size_t &failed_ed) {
    seq.clear();
    failed_ed = 0;
    size_t cur = start;
    for (size_t i = 0; i < vars.size(); ++i) {
        const Var &v = *vars[i];
        uint8_t al = alleles[i];
        if (al == AL_REF) continue;                      // :738-741
        size_t vpos = v.pos;
        if (vpos < cur) { failed_ed += v.alt_ed(); continue; }   // :745-753
        seq.insert(seq.end(), ref + cur, ref + vpos);   // :756-759
        cur = vpos;
        if (al == AL_UNKNOWN) return AVK_ST_BAD_ZYGOSITY;   // :763
        seq.insert(seq.end(), v.a1, v.a1 + v.l1);
        cur += v.ref_len();
    }
    if (cur > end) return AVK_ST_BAD_INPUT;              // get_slice would panic
    seq.insert(seq.end(), ref + cur, ref + end);        // :772-775
    return AVK_ST_OK;
}

// generate_haplotype_sequence(): :685-712.  hap = 0 (Hap1) or 1 (Hap2)
static int generate_haplotype_sequence(const uint8_t *ref, size_t start, size_t end,
                                       const std::vector<const Var *> &vars, const std::vector<uint8_t> &zygs,
                                       int hap, Seq &seq, size_t &failed_ed) {
    std::vector<uint8_t> alleles;
    for (uint8_t z : zygs) {
        if (z == AVK_ZYG_UNKNOWN) return AVK_ST_BAD_ZYGOSITY;
        uint8_t a1, a2;
        zyg_decompose(z, a1, a2);        // same table as :693-700
        alleles.push_back(hap == 0 ? a1 : a2);
    }
    return generate_allele_sequence(ref, start, end, vars, alleles, seq, failed_ed);
}

struct SM4 { uint64_t truth_tp, truth_fn, query_tp, query_fp; };

// perform_basepair_compare(): :611-658 (all values doubled)
static SM4 perform_basepair_compare(const Seq &r, const Seq &t, const Seq &q) {
    uint64_t rt = 2 * (uint64_t)wfa_ed(r.data(), r.size(), t.data(), t.size());
    uint64_t rq = 2 * (uint64_t)wfa_ed(r.data(), r.size(), q.data(), q.size());
    uint64_t tq = 2 * (uint64_t)wfa_ed(t.data(), t.size(), q.data(), q.size());
    uint64_t tp = (rt + rq - tq) / 2;
    return SM4{tp, rt - tp, tp, rq - tp};
}

static std::vector<const Var *> ptrs(const std::vector<Var> &v) {
    std::vector<const Var *> p;
    for (const Var &x : v) p.push_back(&x);
    return p;
}

// add_basepair_stats(): :335-449
static int add_basepair_stats(const uint8_t *ref, size_t start, size_t end,
                              const std::vector<Var> &tv, const std::vector<Var> &qv,
                              Benchmark &b, const OptHaps &oh) {
    Seq ref_seq(ref + start, ref + end);
    std::vector<const Var *> tvp = ptrs(tv), qvp = ptrs(qv);
    for (int hap = 0; hap < 2; ++hap) {
        Seq tseq, qseq;
        size_t ted, qed;
        int st = generate_haplotype_sequence(ref, start, end, tvp, oh.truth_zyg, hap, tseq, ted);
        if (st) return st;
        st = generate_haplotype_sequence(ref, start, end, qvp, oh.query_zyg, hap, qseq, qed);
        if (st) return st;
        // assert_eq!(truth_seq, optimizer truth_seq) :364-367 -- a mismatch is a process panic upstream
        if (tseq != (hap == 0 ? oh.truth_seq1 : oh.truth_seq2)) return AVK_ST_BAD_INPUT;
        if (qseq != (hap == 0 ? oh.query_seq1 : oh.query_seq2)) return AVK_ST_BAD_INPUT;

        SM4 am = perform_basepair_compare(ref_seq, tseq, qseq);
        add4(b.gm.group(-1), AVK_M_BASEPAIR, am.truth_tp, am.truth_fn, am.query_tp, am.query_fp);
        add4(b.gm.group(-1), AVK_M_BASEPAIR, 0, 2 * (uint64_t)ted, 0, 2 * (uint64_t)qed);   // :378-381

        for (int k = 0; k < 8; ++k) {
            const int ft = SUPPORTED_TYPES[k];
            std::vector<const Var *> fv; std::vector<uint8_t> fz;
            for (size_t i = 0; i < qv.size(); ++i) if (qv[i].type == ft) { fv.push_back(&qv[i]); fz.push_back(oh.query_zyg[i]); }
            uint64_t query_tp = 0, query_fp = 0;
            if (!fv.empty()) {                                   // :395-410
                Seq fq; size_t fed;
                st = generate_haplotype_sequence(ref, start, end, fv, fz, hap, fq, fed);
                if (st) return st;
                SM4 fm = perform_basepair_compare(ref_seq, tseq, fq);
                query_tp = fm.query_tp; query_fp = fm.query_fp + 2 * (uint64_t)fed;
            }
            fv.clear(); fz.clear();
            for (size_t i = 0; i < tv.size(); ++i) if (tv[i].type == ft) { fv.push_back(&tv[i]); fz.push_back(oh.truth_zyg[i]); }
            uint64_t truth_tp = 0, truth_fn = 0;
            if (!fv.empty()) {                                   // :422-437
                Seq ftq; size_t fed;
                st = generate_haplotype_sequence(ref, start, end, fv, fz, hap, ftq, fed);
                if (st) return st;
                SM4 fm = perform_basepair_compare(ref_seq, ftq, qseq);
                truth_tp = fm.truth_tp; truth_fn = fm.truth_fn + 2 * (uint64_t)fed;
            }
            add4(b.gm.group(ft), AVK_M_BASEPAIR, truth_tp, truth_fn, query_tp, query_fp);   // :440-444 (always, even all-zero)
        }
    }
    return AVK_ST_OK;
}

// add_record_basepair_stats(): :455-522 (u64 arithmetic wraps like a Rust release build)
static int add_record_basepair_stats(const std::vector<Var> &tv, const std::vector<uint8_t> &tz,
                                     const std::vector<Var> &qv, const std::vector<uint8_t> &qz, Benchmark &b) {
    uint64_t truth_total = 0, query_total = 0;
    uint64_t tt[AVK_N_VARIANT_TYPES] = {0}, qt[AVK_N_VARIANT_TYPES] = {0};
    for (size_t i = 0; i < tv.size(); ++i) {
        uint64_t c = (uint64_t)zyg_count(tz[i]) * tv[i].raw;
        tt[tv[i].type] += c; truth_total += c;
    }
    for (size_t i = 0; i < qv.size(); ++i) {
        uint64_t c = (uint64_t)zyg_count(qz[i]) * qv[i].raw;
        qt[qv[i].type] += c; query_total += c;
    }
    uint64_t *j = b.gm.m[0];
    uint64_t tfn = j[AVK_M_BASEPAIR + 1], qfp = j[AVK_M_BASEPAIR + 3];
    uint64_t ttp = 2 * truth_total - tfn, qtp = 2 * query_total - qfp;
    if (!(ttp >= j[AVK_M_BASEPAIR + 0])) return AVK_ST_TP_UNDERFLOW;   // :492
    if (!(qtp >= j[AVK_M_BASEPAIR + 2])) return AVK_ST_TP_UNDERFLOW;   // :493
    add4(j, AVK_M_RECORD_BP, ttp, tfn, qtp, qfp);
    const uint16_t present = b.gm.mask;                               // :501 clone of the map
    for (int t = 0; t < AVK_N_VARIANT_TYPES; ++t) {
        if (!(present & (1u << t))) continue;
        uint64_t *g = b.gm.m[1 + t];
        uint64_t fn_ = g[AVK_M_BASEPAIR + 1], fp_ = g[AVK_M_BASEPAIR + 3];
        add4(g, AVK_M_RECORD_BP, 2 * tt[t] - fn_, fn_, 2 * qt[t] - fp_, fp_);
    }
    return AVK_ST_OK;
}

// compare_expected_observed(): :296-327
static int compare_expected_observed(size_t ed1, size_t ed2, const std::vector<Var> &vars,
                                     const std::vector<uint8_t> &e1, const std::vector<uint8_t> &o1,
                                     const std::vector<uint8_t> &e2, const std::vector<uint8_t> &o2, Benchmark &b) {
    if (vars.size() != e1.size() || vars.size() != o1.size() || vars.size() != e2.size() || vars.size() != o2.size())
        return AVK_ST_BAD_INPUT;
    b = Benchmark();
    b.ed1 = ed1; b.ed2 = ed2;
    for (size_t i = 0; i < vars.size(); ++i) {
        uint8_t exp = al_count(e1[i]) + al_count(e2[i]);
        uint8_t obs = al_count(o1[i]) + al_count(o2[i]);
        if (!(exp >= obs)) return AVK_ST_TRUTH_FP;     // assert! :322
        int st = b.add_truth_zygosity(vars[i], exp, obs);
        if (st) return st;
    }
    return AVK_ST_OK;
}

// generate_exact_match(): :534-601 (hidden --enable-exact-shortcut path)
static int generate_exact_match(const uint8_t *ref, size_t start, size_t end,
                                const std::vector<Var> &tv, const std::vector<uint8_t> &tz,
                                const std::vector<Var> &qv, const std::vector<uint8_t> &qz,
                                const OptHaps &oh, Benchmark &b) {
    b = Benchmark();
    for (size_t i = 0; i < tv.size(); ++i) {
        if (tz[i] == AVK_ZYG_UNKNOWN) return AVK_ST_BAD_ZYGOSITY;
        uint8_t ev = zyg_count(tz[i]);
        int st = b.add_truth_zygosity(tv[i], ev, ev);
        if (st) return st;
    }
    for (size_t i = 0; i < qv.size(); ++i) {
        if (qz[i] == AVK_ZYG_UNKNOWN) return AVK_ST_BAD_ZYGOSITY;
        uint8_t ev = zyg_count(qz[i]);
        int st = b.add_query_zygosity(qv[i], ev, ev);
        if (st) return st;
    }
    if (oh.truth_seq1 != oh.query_seq1 || oh.truth_seq2 != oh.query_seq2) return AVK_ST_BAD_INPUT;  // assert_eq! :561-562
    uint64_t e1 = wfa_ed(ref + start, end - start, oh.truth_seq1.data(), oh.truth_seq1.size());
    uint64_t e2 = wfa_ed(ref + start, end - start, oh.truth_seq2.data(), oh.truth_seq2.size());
    uint64_t js = 2 * (e1 + e2);
    add4(b.gm.group(-1), AVK_M_BASEPAIR, js, 0, js, 0);
    for (size_t i = 0; i < tv.size(); ++i) {
        uint64_t d = 2 * (uint64_t)tv[i].alt_ed();
        add4(b.gm.group(tv[i].type), AVK_M_BASEPAIR, (uint64_t)zyg_count(tz[i]) * d, 0, 0, 0);
    }
    for (size_t i = 0; i < qv.size(); ++i) {
        uint64_t d = 2 * (uint64_t)qv[i].alt_ed();
        add4(b.gm.group(qv[i].type), AVK_M_BASEPAIR, 0, 0, (uint64_t)zyg_count(qz[i]) * d, 0);
    }
    return AVK_ST_OK;
}

// solve_compare_region(): :122-284
static int solve_compare_region(const uint8_t *ref, size_t ref_len, size_t start, size_t end,
                                const std::vector<Var> &tv, const std::vector<uint8_t> &tz,
                                const std::vector<Var> &qv, const std::vector<uint8_t> &qz,
                                const avk_compare_cfg &cfg, Benchmark &result) {
    if (!(start <= end && end <= ref_len)) return AVK_ST_BAD_INPUT;
    std::vector<OptHaps> all;
    int st = optimize_sequences(ref, start, end, tv, tz, qv, qz, cfg.max_branch_factor, all);
    if (st) return st;

    struct Cand { size_t idx; OptAlleles h1, h2; Benchmark ts, qs; };
    std::vector<Cand> cands;
    for (size_t si = 0; si < all.size(); ++si) {
        const OptHaps &oh = all[si];
        if (cfg.enable_exact_shortcut && oh.is_exact_match()) {        // :171-199
            st = generate_exact_match(ref, start, end, tv, tz, qv, qz, oh, result);
            if (st) return st;
            if (cfg.enable_sequences) {
                result.has_seqs = true;
                result.ref_seq.assign(ref + start, ref + end);
                result.truth_seq1 = oh.truth_seq1; result.truth_seq2 = oh.truth_seq2;
                result.query_seq1 = oh.query_seq1; result.query_seq2 = oh.query_seq2;
            }
            return AVK_ST_OK;
        }
        std::vector<uint8_t> th1, th2, qh1, qh2;                       // :206-211
        for (uint8_t z : oh.truth_zyg) { uint8_t a, b2; zyg_decompose(z, a, b2); th1.push_back(a); th2.push_back(b2); }
        for (uint8_t z : oh.query_zyg) { uint8_t a, b2; zyg_decompose(z, a, b2); qh1.push_back(a); qh2.push_back(b2); }
        Cand c;
        c.idx = si;
        const uint64_t xcap = cfg.exact_gt_max_expansions ? cfg.exact_gt_max_expansions : AVK_EXACT_GT_DEFAULT_MAX_EXPANSIONS;
        st = optimize_gt_alleles(ref, start, end, tv, th1, qv, qh1, c.h1, xcap);   // :214-218
        if (st) return st;
        st = optimize_gt_alleles(ref, start, end, tv, th2, qv, qh2, c.h2, xcap);   // :219-223
        if (st) return st;
        st = compare_expected_observed(oh.ed1, oh.ed2, tv, th1, c.h1.truth_alleles, th2, c.h2.truth_alleles, c.ts);  // :226-235
        if (st) return st;
        if (cfg.enable_sequences) {
            c.ts.has_seqs = true;
            c.ts.ref_seq.assign(ref + start, ref + end);
            c.ts.truth_seq1 = oh.truth_seq1; c.ts.truth_seq2 = oh.truth_seq2;
            c.ts.query_seq1 = oh.query_seq1; c.ts.query_seq2 = oh.query_seq2;
        }
        st = compare_expected_observed(oh.ed1, oh.ed2, qv, qh1, c.h1.query_alleles, qh2, c.h2.query_alleles, c.qs);  // :249-258
        if (st) return st;
        cands.push_back(std::move(c));
    }
    // min_by_key: FIRST minimum (:264-265)
    size_t bi = 0;
    for (size_t i = 1; i < cands.size(); ++i)
        if (cands[i].h1.num_errors + cands[i].h2.num_errors < cands[bi].h1.num_errors + cands[bi].h2.num_errors) bi = i;
    Cand &best = cands[bi];
    result = std::move(best.ts);
    result.add_swap(best.qs);                                          // :269
    st = add_basepair_stats(ref, start, end, tv, qv, result, all[best.idx]);   // :272
    if (st) return st;
    st = add_record_basepair_stats(tv, tz, qv, qz, result);           // :275
    return st;
}

// ---------------------------------------------------------------------------
// merge solver -- src/merge_solver.rs:110-223
// ---------------------------------------------------------------------------
static int64_t variant_delta_length(const std::vector<Var> &v, const std::vector<uint8_t> &z, int &st) {  // :211-223
    int64_t total = 0;
    st = AVK_ST_OK;
    for (size_t i = 0; i < v.size(); ++i) {
        if (z[i] == AVK_ZYG_UNKNOWN) { st = AVK_ST_BAD_ZYGOSITY; return 0; }
        total += ((int64_t)v[i].l1 - (int64_t)v[i].l0) * (int64_t)zyg_count(z[i]);
    }
    return total;
}

static int solve_merge_region(const uint8_t *ref, size_t ref_len, size_t start, size_t end,
                              const std::vector<std::vector<Var>> &vars, const std::vector<std::vector<uint8_t>> &zygs,
                              const avk_merge_cfg &cfg, uint8_t &cls, std::vector<uint8_t> &indices) {
    indices.clear();
    cls = AVK_MERGE_DIFFERENT;
    if (!(start <= end && end <= ref_len)) return AVK_ST_BAD_INPUT;
    const size_t K = vars.size();
    std::vector<int64_t> delta(K);
    for (size_t i = 0; i < K; ++i) { int st; delta[i] = variant_delta_length(vars[i], zygs[i], st); if (st) return st; }
    bool all_identical = true, no_conflict = true;
    std::vector<std::vector<bool>> ms(K, std::vector<bool>(K, false));
    for (size_t i = 0; i < K; ++i) {
        ms[i][i] = true;
        for (size_t j = i + 1; j < K; ++j) {
            bool exact = false;
            if (delta[i] == delta[j]) {                               // :135-143
                std::vector<OptHaps> all;
                int st = optimize_sequences(ref, start, end, vars[i], zygs[i], vars[j], zygs[j], cfg.max_branch_factor, all);
                if (st) return st;
                exact = all[0].is_exact_match();
            }
            all_identical = all_identical && exact;
            no_conflict = no_conflict && (vars[i].empty() || vars[j].empty() || exact);   // :155-157
            if (exact) { ms[i][j] = true; ms[j][i] = true; }
        }
    }
    const size_t maj = K / 2 + 1;                                     // :167
    std::vector<uint8_t> first_maj;
    for (size_t i = 0; i < K; ++i) {
        size_t c = 0;
        for (size_t j = 0; j < K; ++j) c += ms[i][j];
        if (c >= maj) { for (size_t j = 0; j < K; ++j) if (ms[i][j]) first_maj.push_back((uint8_t)j); break; }
    }
    if (all_identical) cls = AVK_MERGE_BASEPAIR_IDENTICAL;            // :174-197
    else if (cfg.no_conflict_enabled && no_conflict) {
        cls = AVK_MERGE_NO_CONFLICT;
        for (size_t i = 0; i < K; ++i) if (!vars[i].empty()) indices.push_back((uint8_t)i);
    } else if (cfg.majority_voting_enabled && !first_maj.empty()) {
        cls = AVK_MERGE_MAJORITY_AGREE; indices = first_maj;
    } else if (cfg.conflict_selection >= 0) {
        cls = AVK_MERGE_CONFLICT_SELECTION; indices.push_back((uint8_t)cfg.conflict_selection);
    } else cls = AVK_MERGE_DIFFERENT;
    return AVK_ST_OK;
}

// ---------------------------------------------------------------------------
// batch plumbing
// ---------------------------------------------------------------------------
static void load_list(const avk_region_batch *b, uint64_t r, uint32_t k, std::vector<Var> &v, std::vector<uint8_t> &z) {
    const avk_variant_table &t = b->variants;
    uint64_t lo = b->var_off[r * b->n_inputs + k], hi = b->var_off[r * b->n_inputs + k + 1];
    v.clear(); z.clear();
    for (uint64_t i = lo; i < hi; ++i) {
        Var x;
        x.pos = t.position[i]; x.type = t.variant_type[i];
        x.a0 = t.allele_pool + t.allele_off[i]; x.l0 = t.a0_len[i];
        x.a1 = x.a0 + x.l0; x.l1 = t.a1_len[i];
        x.raw = t.raw_allele_space[i];
        v.push_back(x);
        z.push_back(t.zygosity[i]);
    }
}

// Inputs the reference never produces (its region builder guarantees them) are
// rejected identically by the oracle and the CUDA library: empty alleles, unknown enum codes, a variant
// outside its window, a list that is not position-sorted.
static bool list_valid(const std::vector<Var> &v, const std::vector<uint8_t> &z, uint64_t start, uint64_t end) {
    for (size_t i = 0; i < v.size(); ++i) {
        if (v[i].l0 == 0 || v[i].l1 == 0) return false;
        if (v[i].type >= AVK_N_VARIANT_TYPES) return false;
        if (z[i] > AVK_ZYG_HOM_ALT) return false;
        if (v[i].pos < start || (uint64_t)v[i].pos + v[i].l0 > end) return false;
        if (i > 0 && v[i - 1].pos > v[i].pos) return false;
    }
    return true;
}

}  // namespace orc

using namespace orc;

extern "C" {

int orc_compare_batch(const avk_region_batch *b, const uint8_t *const *contigs, const uint64_t *contig_lens,
                      uint32_t n_contigs, const avk_compare_cfg *cfg, avk_compare_out *out, int n_threads,
                      avk_work_counters *work) {
    if (!b || !cfg || !out || !out->status || b->n_inputs != 2) return AVK_ERR_INVALID;
    const int64_t n = (int64_t)b->n_regions;
    GroupTypeMetrics totals;
    uint64_t solved = 0, errors = 0;
    std::vector<GroupTypeMetrics> strat(out->strat_totals ? out->n_strata : 0);
    Work total_work;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    (void)n_threads;
#pragma omp parallel
    {
        Work w;
        tl_work = work ? &w : nullptr;
        GroupTypeMetrics loc;
        uint64_t lsolved = 0, lerr = 0;
        std::vector<Var> tv, qv;
        std::vector<uint8_t> tz, qz;
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n; ++r) {
            load_list(b, r, 0, tv, tz);
            load_list(b, r, 1, qv, qz);
            Benchmark res;
            int st;
            uint32_t c = b->contig[r];
            if (c >= n_contigs || b->end[r] > 0x7fff0000u || !list_valid(tv, tz, b->start[r], b->end[r]) || !list_valid(qv, qz, b->start[r], b->end[r])) st = AVK_ST_BAD_INPUT;
            else st = solve_compare_region(contigs[c], contig_lens[c], b->start[r], b->end[r], tv, tz, qv, qz, *cfg, res);
            out->status[r] = st;
            const uint64_t t0 = b->var_off[r * 2], q0 = b->var_off[r * 2 + 1], q1 = b->var_off[r * 2 + 2];
            if (st != AVK_ST_OK) {
                lerr += 1;
                if (out->ed1) out->ed1[r] = 0;
                if (out->ed2) out->ed2[r] = 0;
                if (out->region_metrics) std::memset(out->region_metrics + (size_t)r * AVK_N_GROUPS * AVK_N_METRICS, 0, sizeof(uint64_t) * AVK_N_GROUPS * AVK_N_METRICS);
                if (out->type_mask) out->type_mask[r] = 0;
                for (uint64_t i = t0; i < q1; ++i) {
                    if (out->var_expected) out->var_expected[i] = 0;
                    if (out->var_observed) out->var_observed[i] = 0;
                    if (out->var_class) out->var_class[i] = AVK_CLASS_UNKNOWN;
                }
                if (out->seq_len) for (int s = 0; s < 5; ++s) out->seq_len[r * 5 + s] = 0;
                continue;
            }
            lsolved += 1;
            if (out->ed1) out->ed1[r] = (uint32_t)res.ed1;
            if (out->ed2) out->ed2[r] = (uint32_t)res.ed2;
            if (out->region_metrics) std::memcpy(out->region_metrics + (size_t)r * AVK_N_GROUPS * AVK_N_METRICS, res.gm.m, sizeof(res.gm.m));
            if (out->type_mask) out->type_mask[r] = res.gm.mask;
            for (uint64_t i = t0; i < q0; ++i) {
                const VarMetric &vm = res.truth_data[i - t0];
                if (out->var_expected) out->var_expected[i] = vm.expected;
                if (out->var_observed) out->var_observed[i] = vm.observed;
                if (out->var_class) out->var_class[i] = vm.cls;
            }
            for (uint64_t i = q0; i < q1; ++i) {
                const VarMetric &vm = res.query_data[i - q0];
                if (out->var_expected) out->var_expected[i] = vm.expected;
                if (out->var_observed) out->var_observed[i] = vm.observed;
                if (out->var_class) out->var_class[i] = vm.cls;
            }
            if (out->seq_off && out->seq_len && out->seq_pool) {
                const Seq *ss[5] = {&res.ref_seq, &res.truth_seq1, &res.truth_seq2, &res.query_seq1, &res.query_seq2};
                for (int s = 0; s < 5; ++s) {
                    uint32_t len = res.has_seqs ? (uint32_t)ss[s]->size() : 0;
                    out->seq_len[r * 5 + s] = len;
                    if (len) std::memcpy(out->seq_pool + out->seq_off[r * 5 + s], ss[s]->data(), len);
                }
            }
            for (int g = 0; g < AVK_N_GROUPS; ++g)
                for (int m = 0; m < AVK_N_METRICS; ++m) loc.m[g][m] += res.gm.m[g][m];
            loc.mask |= res.gm.mask;
            if (out->strat_totals && out->strat_off) {
#pragma omp critical(orc_strat)
                for (uint64_t s = out->strat_off[r]; s < out->strat_off[r + 1]; ++s) {
                    GroupTypeMetrics &sg = strat[out->strat_idx[s]];
                    for (int g = 0; g < AVK_N_GROUPS; ++g)
                        for (int m = 0; m < AVK_N_METRICS; ++m) sg.m[g][m] += res.gm.m[g][m];
                }
            }
        }
#pragma omp critical(orc_totals)
        {
            for (int g = 0; g < AVK_N_GROUPS; ++g)
                for (int m = 0; m < AVK_N_METRICS; ++m) totals.m[g][m] += loc.m[g][m];
            totals.mask |= loc.mask;
            solved += lsolved; errors += lerr;
            total_work.add(w);
        }
        tl_work = nullptr;
    }
    if (out->totals) std::memcpy(out->totals, totals.m, sizeof(totals.m));
    if (out->totals_mask) *out->totals_mask = totals.mask;
    if (out->solved_blocks) *out->solved_blocks = solved;
    if (out->error_blocks) *out->error_blocks = errors;
    if (out->strat_totals)
        for (uint32_t s = 0; s < out->n_strata; ++s)
            std::memcpy(out->strat_totals + (size_t)s * AVK_N_GROUPS * AVK_N_METRICS, strat[s].m, sizeof(strat[s].m));
    if (work) {
        work->alignments = total_work.alignments; work->cells = total_work.cells; work->matched_bases = total_work.matched;
        work->search_pops = total_work.search_pops; work->exact_pops = total_work.exact_pops;
        work->alg_bytes = total_work.alg_bytes;
    }
    return AVK_OK;
}

// containment_regions of solve_compare_region (waffle_solver.rs:151-166): var_coordinates() (compare_region.rs:63-74) is
// [min(first truth pos, first query pos), max(LAST truth variant's end, LAST query variant's end)) -- the end comes from
// the last variant of each list, not from the furthest-reaching one -- queried 0-based inclusive as (start, end - 1);
// a stratum contains it when one of its intervals on that contig has first <= start && last >= end - 1
// (stratifications.rs:197-210; every interval is tested, like the COITree query callback).
int orc_containments(const avk_region_batch *b, const avk_strat_intervals *st, uint64_t *mask) {
    if (!b || !st || !mask || b->n_inputs != 2 || st->n_strata > 64) return AVK_ERR_INVALID;
    const avk_variant_table &t = b->variants;
    for (uint64_t r = 0; r < b->n_regions; ++r) {
        mask[r] = 0;
        const uint64_t t0 = b->var_off[r * 2], q0 = b->var_off[r * 2 + 1], q1 = b->var_off[r * 2 + 2];
        uint64_t start = UINT64_MAX, end = 0;
        if (q0 > t0) { start = std::min<uint64_t>(start, t.position[t0]); end = std::max<uint64_t>(end, (uint64_t)t.position[q0 - 1] + t.a0_len[q0 - 1]); }
        if (q1 > q0) { start = std::min<uint64_t>(start, t.position[q0]); end = std::max<uint64_t>(end, (uint64_t)t.position[q1 - 1] + t.a0_len[q1 - 1]); }
        if (!(start < end)) continue;                      // assert!(first < last) :158 would panic; no variants never occurs
        const uint32_t c = b->contig[r];
        if (c >= st->n_contigs) continue;
        const int64_t first = (int64_t)start, last = (int64_t)end - 1;
        for (uint32_t s = 0; s < st->n_strata; ++s) {
            bool included = false;
            for (uint64_t i = st->off[(uint64_t)s * st->n_contigs + c]; i < st->off[(uint64_t)s * st->n_contigs + c + 1]; ++i)
                if ((int64_t)st->first[i] <= first && (int64_t)st->last[i] >= last) included = true;
            if (included) mask[r] |= 1ull << s;
        }
    }
    return AVK_OK;
}

int orc_merge_batch(const avk_region_batch *b, const uint8_t *const *contigs, const uint64_t *contig_lens,
                    uint32_t n_contigs, const avk_merge_cfg *cfg, avk_merge_out *out, int n_threads,
                    avk_work_counters *work) {
    if (!b || !cfg || !out || !out->status || b->n_inputs < 1 || b->n_inputs > 255) return AVK_ERR_INVALID;
    const int64_t n = (int64_t)b->n_regions;
    const uint32_t K = b->n_inputs;
    Work total_work;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    (void)n_threads;
#pragma omp parallel
    {
        Work w;
        tl_work = work ? &w : nullptr;
        std::vector<std::vector<Var>> vars(K);
        std::vector<std::vector<uint8_t>> zygs(K);
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n; ++r) {
            bool ok = b->contig[r] < n_contigs && b->end[r] <= 0x7fff0000u;
            for (uint32_t k = 0; k < K; ++k) { load_list(b, r, k, vars[k], zygs[k]); ok = ok && list_valid(vars[k], zygs[k], b->start[r], b->end[r]); }
            uint8_t cls = AVK_MERGE_DIFFERENT;
            std::vector<uint8_t> idx;
            int st = ok ? solve_merge_region(contigs[b->contig[r]], contig_lens[b->contig[r]], b->start[r], b->end[r], vars, zygs, *cfg, cls, idx)
                        : AVK_ST_BAD_INPUT;
            out->status[r] = st;
            if (st != AVK_ST_OK) { cls = AVK_MERGE_DIFFERENT; idx.clear(); }
            if (out->classification) out->classification[r] = cls;
            if (out->n_indices) out->n_indices[r] = (uint8_t)idx.size();
            if (out->indices) {
                for (uint32_t k = 0; k < K; ++k) out->indices[(size_t)r * K + k] = k < idx.size() ? idx[k] : 0xFF;
            }
        }
#pragma omp critical(orc_mtotals)
        total_work.add(w);
        tl_work = nullptr;
    }
    if (work) {
        work->alignments = total_work.alignments; work->cells = total_work.cells; work->matched_bases = total_work.matched;
        work->search_pops = total_work.search_pops; work->exact_pops = total_work.exact_pops;
        work->alg_bytes = total_work.alg_bytes;
    }
    return AVK_OK;
}

// ----- fine-grained hooks used by tests/test_oracle_golden.py -----------------

uint64_t orc_wfa_ed(const uint8_t *a, uint64_t la, const uint8_t *b, uint64_t lb) { return wfa_ed(a, la, b, lb); }
uint64_t orc_edit_distance(const uint8_t *a, uint64_t la, const uint8_t *b, uint64_t lb) { return edit_distance(a, la, b, lb); }

void *orc_dwfa_new(uint64_t max_ed) { DWFALite *d = new DWFALite(); d->max_ed = max_ed == UINT64_MAX ? SIZE_MAX : (size_t)max_ed; return d; }
void orc_dwfa_free(void *p) { delete (DWFALite *)p; }
void *orc_dwfa_clone(void *p) { return new DWFALite(*(DWFALite *)p); }
int orc_dwfa_update(void *p, const uint8_t *b, uint64_t lb, const uint8_t *o, uint64_t lo) { return (int)((DWFALite *)p)->update(b, lb, o, lo); }
int orc_dwfa_finalize(void *p, const uint8_t *b, uint64_t lb, const uint8_t *o, uint64_t lo) { return (int)((DWFALite *)p)->finalize(b, lb, o, lo); }
uint64_t orc_dwfa_ed(void *p) { return ((DWFALite *)p)->ed; }
uint64_t orc_dwfa_wavefront(void *p, uint64_t *out, uint64_t cap) {
    DWFALite *d = (DWFALite *)p;
    for (size_t i = 0; i < d->wf.size() && i < cap; ++i) out[i] = d->wf[i];
    return d->wf.size();
}
int orc_dwfa_equal(void *a, void *b) {
    DWFALite *x = (DWFALite *)a, *y = (DWFALite *)b;
    return x->ed == y->ed && x->wf == y->wf && x->finalized == y->finalized && x->max_ed == y->max_ed;
}

static Var mkvar(uint32_t pos, const uint8_t *a0, uint32_t l0, const uint8_t *a1, uint32_t l1) {
    Var v; v.pos = pos; v.type = 0; v.a0 = a0; v.l0 = l0; v.a1 = a1; v.l1 = l1; v.raw = std::max(l0, l1); return v;
}
void *orc_hap_new(uint64_t start, uint64_t max_ed) { return new HapDWFA(start, max_ed == UINT64_MAX ? SIZE_MAX : (size_t)max_ed); }
void orc_hap_free(void *p) { delete (HapDWFA *)p; }
// allele: 1 = REF, 2 = ALT; sync < 0 => None.  returns success (0/1), -1 unknown allele, -2 DWFA error
int orc_hap_extend(void *p, const uint8_t *ref, int is_truth, uint32_t pos, const uint8_t *a0, uint32_t l0,
                   const uint8_t *a1, uint32_t l1, int allele, int64_t sync) {
    DErr de;
    Var v = mkvar(pos, a0, l0, a1, l1);
    int s = ((HapDWFA *)p)->extend_variant(ref, is_truth != 0, v, (uint8_t)allele, sync >= 0, sync >= 0 ? (size_t)sync : 0, &de);
    if (s < 0) return -1;
    if (de != DErr::Ok) return -2;
    return s;
}
int orc_hap_finalize(void *p, const uint8_t *ref, uint64_t region_end) { return (int)((HapDWFA *)p)->finalize_dwfa(ref, region_end); }
uint64_t orc_hap_ed(void *p) { return ((HapDWFA *)p)->dwfa.ed; }
uint64_t orc_hap_skip(void *p) { HapDWFA *h = (HapDWFA *)p; return h->truth.skip + h->query.skip; }
uint64_t orc_hap_cost(void *p) { return ((HapDWFA *)p)->total_cost(); }
uint64_t orc_hap_seq(void *p, int is_truth, uint8_t *out, uint64_t cap) {
    HapDWFA *h = (HapDWFA *)p;
    const Seq &s = is_truth ? h->truth.seq : h->query.seq;
    std::memcpy(out, s.data(), std::min<uint64_t>(cap, s.size()));
    return s.size();
}
uint64_t orc_hap_alleles(void *p, int is_truth, uint8_t *out, uint64_t cap) {
    HapDWFA *h = (HapDWFA *)p;
    const std::vector<uint8_t> &s = is_truth ? h->truth.alleles : h->query.alleles;
    std::memcpy(out, s.data(), std::min<uint64_t>(cap, s.size()));
    return s.size();
}

// optimize_sequences on region 0 of a 2-input batch.  Result i is written as a
// record: zygosities (nt + nq bytes) then 10 u64 {ed1, ed2, tvs1, tvs2, qvs1, qvs2, len ts1, ts2, qs1, qs2}
// and the four sequences concatenated into seq_out at stride seq_stride.
int orc_optimize_sequences(const avk_region_batch *b, const uint8_t *ref, uint64_t ref_len, uint32_t max_branch_factor,
                           uint32_t max_results, uint32_t *n_results, uint8_t *zyg_out, uint64_t *num_out,
                           uint8_t *seq_out, uint64_t seq_stride) {
    (void)ref_len;
    std::vector<Var> tv, qv; std::vector<uint8_t> tz, qz;
    load_list(b, 0, 0, tv, tz); load_list(b, 0, 1, qv, qz);
    std::vector<OptHaps> all;
    int st = optimize_sequences(ref, b->start[0], b->end[0], tv, tz, qv, qz, max_branch_factor, all);
    *n_results = (uint32_t)all.size();
    if (st) return st;
    const size_t nv = tv.size() + qv.size();
    for (size_t i = 0; i < all.size() && i < max_results; ++i) {
        const OptHaps &o = all[i];
        std::memcpy(zyg_out + i * nv, o.truth_zyg.data(), tv.size());
        std::memcpy(zyg_out + i * nv + tv.size(), o.query_zyg.data(), qv.size());
        uint64_t *n = num_out + i * 10;
        n[0] = o.ed1; n[1] = o.ed2; n[2] = o.truth_vs1; n[3] = o.truth_vs2; n[4] = o.query_vs1; n[5] = o.query_vs2;
        n[6] = o.truth_seq1.size(); n[7] = o.truth_seq2.size(); n[8] = o.query_seq1.size(); n[9] = o.query_seq2.size();
        const Seq *ss[4] = {&o.truth_seq1, &o.truth_seq2, &o.query_seq1, &o.query_seq2};
        for (int s = 0; s < 4; ++s) std::memcpy(seq_out + (i * 4 + s) * seq_stride, ss[s]->data(), std::min<size_t>(seq_stride, ss[s]->size()));
    }
    return AVK_ST_OK;
}

// optimize_gt_alleles on region 0; the batch `zygosity` column carries Allele codes (1 REF / 2 ALT).
int orc_optimize_gt_alleles(const avk_region_batch *b, const uint8_t *ref, uint8_t *truth_out, uint8_t *query_out, uint64_t *num_errors) {
    std::vector<Var> tv, qv; std::vector<uint8_t> ta, qa;
    load_list(b, 0, 0, tv, ta); load_list(b, 0, 1, qv, qa);
    OptAlleles oa;
    int st = optimize_gt_alleles(ref, b->start[0], b->end[0], tv, ta, qv, qa, oa);
    if (st) return st;
    std::memcpy(truth_out, oa.truth_alleles.data(), oa.truth_alleles.size());
    std::memcpy(query_out, oa.query_alleles.data(), oa.query_alleles.size());
    *num_errors = oa.num_errors;
    return AVK_ST_OK;
}

// generate_haplotype_sequence on list (0, 0) of the batch.
int orc_generate_haplotype_sequence(const avk_region_batch *b, const uint8_t *ref, int hap, uint8_t *out, uint64_t cap,
                                    uint64_t *len, uint64_t *failed_ed) {
    std::vector<Var> v; std::vector<uint8_t> z;
    load_list(b, 0, 0, v, z);
    Seq s; size_t fe;
    int st = generate_haplotype_sequence(ref, b->start[0], b->end[0], ptrs(v), z, hap, s, fe);
    if (st) return st;
    std::memcpy(out, s.data(), std::min<uint64_t>(cap, s.size()));
    *len = s.size(); *failed_ed = fe;
    return AVK_ST_OK;
}

void orc_perform_basepair_compare(const uint8_t *r, uint64_t lr, const uint8_t *t, uint64_t lt, const uint8_t *q, uint64_t lq, uint64_t *out4) {
    SM4 m = perform_basepair_compare(Seq(r, r + lr), Seq(t, t + lt), Seq(q, q + lq));
    out4[0] = m.truth_tp; out4[1] = m.truth_fn; out4[2] = m.query_tp; out4[3] = m.query_fp;
}

int64_t orc_variant_delta_length(const avk_region_batch *b, uint32_t k) {
    std::vector<Var> v; std::vector<uint8_t> z;
    load_list(b, 0, k, v, z);
    int st;
    return variant_delta_length(v, z, st);
}

// ---------------------------------------------------------------------------
// Region builder -- src/parsing/region_generation.rs:352-469 for one contig and one BED interval spanning it
// (SURVEY 8f N1).  The reference has no tests for this function (region_generation.rs:814-821): parity unpinned;
// it is pinned here against the generator's host builder (aardvark_b200/synth.py::cluster_regions) instead.
// `out` arrays are caller-allocated with room for every input variant; returns 0 and fills out->n_regions etc.
// ---------------------------------------------------------------------------
int orc_build_regions(const avk_callsets *in, uint64_t contig_len, uint32_t contig, uint32_t flank, uint64_t first_region_id,
                      avk_region_batch *out) {
    const avk_variant_table &t = in->variants;
    const uint32_t K = in->n_inputs;
    std::vector<uint32_t> order;                                  // all inputs concatenated in input order (:352-366)
    order.reserve(t.n_variants);
    for (uint64_t i = 0; i < t.n_variants; ++i)
        if ((uint64_t)t.position[i] + t.a0_len[i] <= contig_len) order.push_back((uint32_t)i);   // fully contained (:551)
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return t.position[a] < t.position[b]; });   // :367-373
    std::vector<uint32_t> input_of(t.n_variants);
    for (uint32_t k = 0; k < K; ++k) for (uint64_t i = in->input_off[k]; i < in->input_off[k + 1]; ++i) input_of[i] = k;
    avk_variant_table &o = const_cast<avk_variant_table &>(out->variants);
    uint64_t n = 0, nv = 0, pool = 0;
    uint64_t *region_id = const_cast<uint64_t *>(out->region_id), *var_off = const_cast<uint64_t *>(out->var_off);
    uint32_t *rc = const_cast<uint32_t *>(out->contig), *rs = const_cast<uint32_t *>(out->start), *re = const_cast<uint32_t *>(out->end);
    std::vector<std::vector<uint32_t>> cur(K);
    bool open = false;
    uint64_t w_start = 0, w_end = 0;
    var_off[0] = 0;
    auto flush = [&]() {                                          // :403-409, :459-466
        region_id[n] = first_region_id + n; rc[n] = contig; rs[n] = (uint32_t)w_start; re[n] = (uint32_t)w_end;
        for (uint32_t k = 0; k < K; ++k) {
            for (uint32_t i : cur[k]) {
                const_cast<uint32_t *>(o.position)[nv] = t.position[i]; const_cast<uint8_t *>(o.variant_type)[nv] = t.variant_type[i];
                const_cast<uint8_t *>(o.zygosity)[nv] = t.zygosity[i]; const_cast<uint32_t *>(o.raw_allele_space)[nv] = t.raw_allele_space[i];
                const_cast<uint32_t *>(o.allele_off)[nv] = (uint32_t)pool; const_cast<uint32_t *>(o.a0_len)[nv] = t.a0_len[i];
                const_cast<uint32_t *>(o.a1_len)[nv] = t.a1_len[i];
                std::memcpy(const_cast<uint8_t *>(o.allele_pool) + pool, t.allele_pool + t.allele_off[i], (size_t)t.a0_len[i] + t.a1_len[i]);
                pool += (uint64_t)t.a0_len[i] + t.a1_len[i];
                nv += 1;
            }
            var_off[n * K + k + 1] = nv;
            cur[k].clear();
        }
        n += 1;
    };
    for (uint32_t i : order) {
        const uint64_t p = t.position[i];
        if (open && p >= w_end) { flush(); open = false; }        // :396-409
        const uint64_t vend = std::min<uint64_t>(p + t.a0_len[i] + flank, contig_len);   // :411-429
        if (!open) { w_start = p > flank ? p - flank : 0; w_end = vend; open = true; }
        else w_end = std::max(w_end, vend);
        cur[input_of[i]].push_back(i);
    }
    if (open) flush();                                            // :449-469
    out->n_regions = n; out->n_inputs = K; o.n_variants = nv; o.allele_pool_len = pool;
    return 0;
}

// RegionIterator::next over every contig and BED interval (region_generation.rs:276-479), literally: per contig the variants
// fully inside the contig's full region [first interval start, last interval end) (:551, is_variant_contained :764-777) are
// concatenated in input order, stable-sorted by position and consumed from a deque interval by interval.
int orc_build_regions_bed(const avk_callsets *in, const uint32_t *variant_contig, const avk_bed_intervals *bed, const uint64_t *contig_lens,
                          uint32_t n_contigs, uint32_t flank, uint64_t first_region_id, avk_region_batch *out) {
    const avk_variant_table &t = in->variants;
    const uint32_t K = in->n_inputs;
    std::vector<uint32_t> input_of(t.n_variants);
    for (uint32_t k = 0; k < K; ++k) for (uint64_t i = in->input_off[k]; i < in->input_off[k + 1]; ++i) input_of[i] = k;
    avk_variant_table &o = const_cast<avk_variant_table &>(out->variants);
    uint64_t n = 0, nv = 0, pool = 0;
    uint64_t *region_id = const_cast<uint64_t *>(out->region_id), *var_off = const_cast<uint64_t *>(out->var_off);
    uint32_t *rc = const_cast<uint32_t *>(out->contig), *rs = const_cast<uint32_t *>(out->start), *re = const_cast<uint32_t *>(out->end);
    var_off[0] = 0;
    for (uint32_t c = 0; c < n_contigs; ++c) {
        std::vector<std::pair<uint64_t, uint64_t>> ivs;                       // 0-based half-open
        if (bed) for (uint64_t j = bed->first[c]; j < bed->first[c + 1]; ++j) ivs.push_back({bed->start[j], bed->end[j]});
        else ivs.push_back({0, contig_lens[c]});
        if (ivs.empty()) continue;                                            // contigs without intervals are never visited (:283-284)
        const uint64_t full_start = ivs.front().first, full_end = ivs.back().second, chrom_length = contig_lens[c];
        std::vector<uint32_t> order;
        for (uint64_t i = 0; i < t.n_variants; ++i) {
            if (variant_contig[i] != c) continue;
            const uint64_t vs = t.position[i], last = vs + t.a0_len[i] - 1;  // is_variant_contained :764-777
            if (vs >= full_start && vs < full_end && last >= full_start && last < full_end && vs + t.a0_len[i] <= chrom_length) order.push_back((uint32_t)i);
        }
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return t.position[a] < t.position[b]; });   // :367-373
        size_t front = 0;                                                     // the deque's front
        for (const auto &iv : ivs) {
            std::vector<std::vector<uint32_t>> cur(K);
            bool have_start = false, have_end = false;
            uint64_t w_start = 0, w_end = 0;
            auto flush = [&]() {
                region_id[n] = first_region_id + n; rc[n] = c; rs[n] = (uint32_t)w_start; re[n] = (uint32_t)w_end;
                for (uint32_t k = 0; k < K; ++k) {
                    for (uint32_t i : cur[k]) {
                        const_cast<uint32_t *>(o.position)[nv] = t.position[i]; const_cast<uint8_t *>(o.variant_type)[nv] = t.variant_type[i];
                        const_cast<uint8_t *>(o.zygosity)[nv] = t.zygosity[i]; const_cast<uint32_t *>(o.raw_allele_space)[nv] = t.raw_allele_space[i];
                        const_cast<uint32_t *>(o.allele_off)[nv] = (uint32_t)pool; const_cast<uint32_t *>(o.a0_len)[nv] = t.a0_len[i];
                        const_cast<uint32_t *>(o.a1_len)[nv] = t.a1_len[i];
                        std::memcpy(const_cast<uint8_t *>(o.allele_pool) + pool, t.allele_pool + t.allele_off[i], (size_t)t.a0_len[i] + t.a1_len[i]);
                        pool += (uint64_t)t.a0_len[i] + t.a1_len[i];
                        nv += 1;
                    }
                    var_off[n * K + k + 1] = nv;
                    cur[k].clear();
                }
                n += 1;
            };
            while (front < order.size()) {
                const uint32_t i = order[front];
                const uint64_t vs = t.position[i], ve = vs + t.a0_len[i];
                if (vs < iv.first) { front += 1; continue; }                  // Before: skipped (:388-391)
                if (vs >= iv.second) break;                                   // After: left for the next interval (:432-437)
                front += 1;
                if (ve > iv.second) continue;                                 // Overlapping: consumed, unused (:438-442)
                if (have_end && vs >= w_end) {                                // :396-418 (window_end is deliberately not reset)
                    flush();
                    have_start = false;
                }
                if (!have_start) { w_start = vs > flank ? vs - flank : 0; have_start = true; }
                const uint64_t vfe = std::min<uint64_t>(vs + t.a0_len[i] + flank, chrom_length);
                w_end = have_end ? std::max(w_end, vfe) : vfe;
                have_end = true;
                cur[input_of[i]].push_back(i);
            }
            if (have_start && have_end) flush();                              // :447-466
        }
    }
    out->n_regions = n; out->n_inputs = K; o.n_variants = nv; o.allele_pool_len = pool;
    return 0;
}

// parse_variant / parse_genotype / get_variant_type (region_generation.rs:565-758) over inflated VCF record lines, written with
// std::string splitting (independent of the device parser in aardvark_b200/csrc/avk_vcf.cuh).  Returns 0, or 1 + the index of
// the first line the reference would fail on (*err_code: 1 columns, 2 POS, 3 no GT key, 4 GT, 5 ALT index, 6 SVTYPE, 7 contig,
// 8 empty allele, 9 Variant constructor).
static std::vector<std::string> split_str(const std::string &s, char d) {
    std::vector<std::string> out;
    size_t b = 0;
    for (;;) { const size_t e = s.find(d, b); if (e == std::string::npos) { out.push_back(s.substr(b)); break; } out.push_back(s.substr(b, e - b)); b = e + 1; }
    return out;
}
int orc_vcf_parse(const uint8_t *text, uint64_t len, const char *const *contig_names, uint32_t n_contigs, uint32_t sample_index, int enable_trimming,
                  avk_vcf_out *out, int32_t *err_code) {
    uint64_t nv = 0, pool = 0, line_no = 0;
    size_t b = 0;
    const std::string all((const char *)text, (size_t)len);
    auto fail = [&](int code) { if (err_code) *err_code = code; return (int)(1 + line_no); };
    while (b < all.size()) {
        size_t e = all.find('\n', b);
        if (e == std::string::npos) e = all.size();
        std::string line = all.substr(b, e - b);
        b = e + 1;
        while (!line.empty() && line.back() == '\r') line.pop_back();
        const uint64_t this_line = line_no;
        (void)this_line;
        if (line.empty() || line[0] == '#') { line_no += 1; continue; }
        const std::vector<std::string> col = split_str(line, '\t');
        if (col.size() < 10 || col.size() - 9 <= sample_index) return fail(1);
        uint32_t contig = n_contigs;
        for (uint32_t c = 0; c < n_contigs; ++c) if (col[0] == contig_names[c]) { contig = c; break; }
        if (contig == n_contigs) return fail(7);
        if (col[1].empty() || col[1].find_first_not_of("0123456789") != std::string::npos || col[1].size() > 10) return fail(2);
        const uint64_t pos1 = std::stoull(col[1]);
        if (pos1 == 0 || pos1 > 0xffffffffull) return fail(2);
        const std::vector<std::string> fmt = split_str(col[8], ':'), smp = split_str(col[9 + sample_index], ':');
        int gi = -1;
        for (size_t k = 0; k < fmt.size(); ++k) if (fmt[k] == "GT") { gi = (int)k; break; }
        if (gi < 0) return fail(3);                                             // "Missing GT" (:579-580)
        if ((size_t)gi >= smp.size() || smp[gi].empty() || smp[gi] == ".") { line_no += 1; continue; }   // GT = '.': no-op (:581-587)
        // parse_genotype (:660-712)
        std::string g = smp[gi];
        if (g[0] == '/' || g[0] == '|') g = g.substr(1);
        std::vector<int> al;
        bool phased = false;
        size_t i = 0;
        for (;;) {
            if (al.size() == 2 || i >= g.size()) return fail(4);
            if (g[i] == '.') { al.push_back(0); i += 1; }
            else {
                if (!isdigit((unsigned char)g[i])) return fail(4);
                int v = 0;
                while (i < g.size() && isdigit((unsigned char)g[i])) { v = v * 10 + (g[i] - '0'); if (v > 60000) return fail(4); i += 1; }
                al.push_back(v);
            }
            if (i == g.size()) break;
            if (g[i] == '|') phased = true; else if (g[i] != '/') return fail(4);
            i += 1;
        }
        if (al.size() == 1) al.push_back(al[0]);                                 // hemizygous treated as homozygous (:671-673)
        std::vector<std::pair<int, int>> picks;                                  // (alt index, zygosity)
        if (al[0] == al[1]) { if (al[0] != 0) picks.push_back({al[0], AVK_ZYG_HOM_ALT}); }
        else {
            if (al[0] != 0) picks.push_back({al[0], phased ? AVK_ZYG_PHASED_HET10 : AVK_ZYG_UNPHASED_HET});
            if (al[1] != 0) picks.push_back({al[1], phased ? AVK_ZYG_PHASED_HET01 : AVK_ZYG_UNPHASED_HET});
        }
        if (picks.empty()) { line_no += 1; continue; }
        // INFO tags (get_variant_type :723-746)
        int sv = -1; bool trid = false, bad_sv = false;
        for (const std::string &f : split_str(col[7], ';')) {
            if (f.rfind("SVTYPE=", 0) == 0 && f.size() > 7) {
                const std::string v = f.substr(7);
                if (v == "BND") sv = AVK_VT_SV_BREAKEND; else if (v == "DEL") sv = AVK_VT_SV_DELETION; else if (v == "DUP") sv = AVK_VT_SV_DUPLICATION;
                else if (v == "INS") sv = AVK_VT_SV_INSERTION; else bad_sv = true;
            }
            if (f.rfind("TRID=", 0) == 0 && f.size() > 5) trid = true;
        }
        const std::vector<std::string> alts = split_str(col[4], ',');
        for (const auto &pk : picks) {
            if ((size_t)pk.first > alts.size()) return fail(5);
            std::string r = col[3], a = alts[pk.first - 1];
            if (r.empty() || a.empty()) return fail(8);
            if (a == "*") continue;                                             // :596-599
            if (a[0] == '<') continue;                                          // :603-606
            const size_t raw = std::max(r.size(), a.size());                    // :610-612
            while (enable_trimming && r.size() > 1 && a.size() > 1 && r.back() == a.back()) { r.pop_back(); a.pop_back(); }   // :615-618
            if (r.size() > 10000 || a.size() > 10000) continue;                 // :621-626
            if (bad_sv) return fail(6);
            int vt;
            if (sv >= 0) vt = sv;
            else if (trid) vt = a.size() < r.size() ? AVK_VT_TR_CONTRACTION : AVK_VT_TR_EXPANSION;
            else vt = (r.size() == 1 && a.size() == 1) ? AVK_VT_SNV : (r.size() == 1 ? AVK_VT_INSERTION : (a.size() == 1 ? AVK_VT_DELETION : AVK_VT_INDEL));
            if (vt == AVK_VT_SV_BREAKEND || vt == AVK_VT_SV_DUPLICATION) continue;   // :641-644
            if ((vt == AVK_VT_SV_DELETION && (r.size() <= 1 || a.size() > r.size())) || (vt == AVK_VT_SV_INSERTION && a.size() < r.size())) return fail(9);   // variants.rs:232-290
            if (nv >= out->cap_variants || pool + r.size() + a.size() > out->cap_pool) return -1;
            out->contig[nv] = contig; out->position[nv] = (uint32_t)(pos1 - 1); out->variant_type[nv] = (uint8_t)vt; out->zygosity[nv] = (uint8_t)pk.second;
            out->raw_allele_space[nv] = (uint32_t)raw; out->allele_off[nv] = (uint32_t)pool; out->a0_len[nv] = (uint32_t)r.size(); out->a1_len[nv] = (uint32_t)a.size();
            std::memcpy(out->allele_pool + pool, r.data(), r.size()); std::memcpy(out->allele_pool + pool + r.size(), a.data(), a.size());
            pool += r.size() + a.size();
            nv += 1;
        }
        line_no += 1;
    }
    out->n_variants = nv; out->allele_pool_len = pool;
    return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
