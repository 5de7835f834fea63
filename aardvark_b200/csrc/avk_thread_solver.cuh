// avk_thread_solver.cuh -- solve_compare_region for ONE cluster by ONE thread, as a resumable state machine.
//
// The warp solvers of avk_solver.cuh put a whole warp on one cluster; for the common small cluster (a handful of
// variants, windows of a few hundred bases, edit distances of a few units) most of what such a warp executes is
// warp-uniform scalar bookkeeping replicated on 32 lanes.  Here a cluster belongs to one thread -- 32 clusters per
// warp -- and the search is restated so that a thread needs next to no memory (< 900 bytes of shared memory):
//
//   * A search node is NOT stored.  DWFALite is path independent (after every update the state equals the state of
//     one from-scratch update on the current sequences: increase_edit_distance only fires while no diagonal touches
//     an end, dynamic_wfa.rs:77), and a haplotype sequence is a pure function of the alleles assigned so far.  A queued
//     node is therefore 12 bytes: (cost, id) key, the two allele bit masks, depth and the two edit distances.  A popped
//     node rebuilds its trackers by replaying its alleles (HaplotypeTracker, haplotype_dwfa.rs:175-227) and its
//     wavefront is known in closed form when its edit distance is 0 (the single diagonal stands at the end of the
//     shorter sequence) or recomputed from scratch otherwise.
//   * A haplotype sequence is never materialised: it is a list of pieces (reference runs and ALT alleles) and the
//     longest-common-prefix walk reads the bytes in place; two reference runs at the same coordinate match without
//     being read.
//   * The solve is cut into STEPS (one queue pop of optimize_sequences, one queue pop of optimize_gt_alleles, the final
//     scoring) so that the 32 threads of a warp, each on its own cluster, can be kept on the same code: the kernel runs
//     the step kinds in warp-wide rounds (k_compare_thread).
//
// Exactly the reference's searches are replayed -- same priority keys, node ids, quotas, prunes and tie-breaks
// (query_optimizer.rs:166-365, exact_gt_optimizer.rs:108-357, waffle_solver.rs:122-522) -- with the exactness-preserving
// shortcuts of RegionSolver::compare_score.  Whatever does not fit the fixed per-thread workspace
// (more than TS_MAXN variants, queue / result / edit-distance / piece capacity, sequence bundle or exact shortcut
// requested) is REJECTED before any output is written and solved by the warp kernels instead.
//
// The code is plain scalar C++ (AVK_HD): the same source is compiled for the host by tests/ts_host.cpp and checked
// against the CPU oracle there, cluster by cluster, before it ever runs on a GPU.
#pragma once
#include "avk_layout.h"

namespace avk_ts {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
using namespace avk;

enum { TS_MAXN = 12, TS_EDCAP = 16, TS_QCAP = 16, TS_RESCAP = 8, TS_MAXALT = 6, TS_MAXP = 2 * TS_MAXALT + 1, TS_MAXSLOT = 4 };
enum { TS_REJECT = -1 };   // >= 0: AVK_ST_*
enum { TS_WF = 2 * TS_EDCAP + 2 };
enum { PH_FETCH = 0, PH_SEARCH = 1, PH_EXACT = 2, PH_FINISH = 3, PH_DONE = 4 };   // what the thread has to do next

struct VarInfo { u16 pos, aoff; u8 l0, l1, alted, pad; };                   // pos relative to the region start
struct QEnt { u32 key; u16 a1, a2; u8 depth, ed1, ed2, pad; };              // optimize_sequences: key = cost << 16 | id
struct XEnt { u32 key; u16 keep; u8 depth, pad; };                          // optimize_gt_alleles: key = errors << 27 | (31 - good) << 22 | id
struct ResEnt { u16 a1, a2; u8 ed1, ed2, tvs1, tvs2, qvs1, qvs2, p0, p1; };
// pieces of a virtual haplotype sequence, alternating reference run (even index) / ALT allele (odd index)
struct PSeq {
    u16 ls[TS_MAXP + 1];       // logical start of piece k; ls[n] == length
    u16 src[TS_MAXP];          // even k: reference position relative to the region start; odd k: offset into the allele bytes
};
enum { TS_XCAP = (TS_QCAP * sizeof(QEnt)) / sizeof(XEnt) };

// per-thread workspace (shared memory on the GPU)
struct Work {
    VarInfo var[TS_MAXN];
    u8 vtype[TS_MAXN], zyg[TS_MAXN], slot[TS_MAXN];
    u8 bucket[TS_MAXN + 2];
    u16 truth_mask;                 // bit oi: order entry oi is a truth variant
    union {
        QEnt q[TS_QCAP];
        XEnt x[TS_XCAP];
    };
    ResEnt res[TS_RESCAP];
    u16 wf[3][TS_WF];               // parent hap 0 / hap 1, child (also the scoring alignments)
    PSeq seq[3];
    u16 pad;
};
// one workspace per thread, side by side in shared memory: an odd number of 32-bit words per workspace keeps the 32 lanes
// of a warp on 32 different banks when they touch the same field
static_assert(sizeof(Work) % 4 == 0 && (sizeof(Work) / 4) % 2 == 1, "sizeof(Work) must be an odd number of words");

// counters of the work actually executed (same meaning as avk_work_counters)
struct Counters { u64 cells, matched; u32 alignments, spops, xpops; };

// read-only view of one cluster
struct Cluster {
    const u8 *ref;     // region window: base at absolute position start + x is ref[x]
    const u8 *recs;    // N VI_* records in merged order (digest)
    const u8 *alle;    // allele bytes of the digest
    int start, end, N, nT, nQ, mbf, n_slots;
    u32 slot_types;    // 4 bits per slot: variant type of metric-row slot k
};

// scalars of a replayed sequence (its pieces are in a PSeq)
struct SeqInfo { int len, ref_pos, skip, n_alt, last_ok; };

static AVK_HD inline u32 rec32(const Cluster &c, int oi, int field) { return *(const u32 *)(c.recs + (size_t)VI_SIZE * oi + field); }
static AVK_HD inline int min_i(int a, int b) { return a < b ? a : b; }
static AVK_HD inline int max_i(int a, int b) { return a > b ? a : b; }

// What a solved cluster hands to its sink: everything needed to write the outputs of solve_compare_region.
struct Solution {
    u32 ed1, ed2;
    u16 type_mask;
    int n;
    u8 exp[TS_MAXN], obs[TS_MAXN];          // per order entry, already toggled for query entries
    int n_rows;                             // 1 + n_slots
    u8 row_group[1 + TS_MAXSLOT];           // group index of each row (0 joint, 1 + type)
    u64 rows[1 + TS_MAXSLOT][AVK_N_METRICS];
};

// GroupMetrics::add_truth_zygosity (grouped_metrics.rs:183-227) on row g; col 0 = truth columns, 2 = query columns
static AVK_HD inline void gm_add(u64 *g, int col, u64 wgt, int exp, int obs) {
    const int mn = exp < obs ? exp : obs;
    g[AVK_M_HAP + col] += (u64)mn;
    g[AVK_M_WEIGHTED_HAP + col] += (u64)mn * wgt;
    if (exp == obs) {
        g[AVK_M_GT + col] += 1;
    } else {
        g[AVK_M_HAP + col + 1] += (u64)(exp - obs);
        g[AVK_M_WEIGHTED_HAP + col + 1] += (u64)(exp - obs) * wgt;
        g[AVK_M_GT + col + 1] += 1;
        if (obs > 0) g[(col == 0) ? AVK_M_GT_TRUTH_FN_GT : AVK_M_GT_QUERY_FP_GT] += 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// One thread's solver: `c` and the scalar members live in registers, `w` in shared memory.
struct Solver {
    Cluster c;
    Work *wp;
    Counters *ctr;
    int phase;
    int rc;                        // status to report when phase == PH_FINISH and rc != AVK_ST_OK
    // optimize_sequences
    int qn, nres;
    u32 best, next_id;
    // scoring: iteration over the equal-best results and their two haplotypes
    int ri, hi, h, best_r, best_total, total, budget;
    bool lost;
    u32 keep[2], best_keep[2];
    // optimize_gt_alleles
    int xn, best_err, min_sync, af_index, af_counts;
    bool have_best;
    u32 x_next_id, hap_alt, x_keep;

    AVK_HD Work &w() const { return *wp; }
    AVK_HD int sync_pos(int oi) const { return (oi == c.N - 1) ? (c.end - c.start) : (int)w().var[oi + 1].pos; }   // query_optimizer.rs:258-265 (relative)
    AVK_HD bool is_truth(int oi) const { return (w().truth_mask >> oi) & 1; }
    AVK_HD int slot_type(int k) const { return (int)((c.slot_types >> (4 * k)) & 15u); }

    // HaplotypeTracker replay (haplotype_dwfa.rs:175-227): the sequence of `side` (0 truth, 1 query) after the first
    // `depth` order entries with ALT where `mask` has the bit; to_end: copy_reference(region end) at the end.
    // Positions are relative to the region start.  returns false when the sequence has more than TS_MAXALT spliced ALTs.
    template <bool PIECES>
    AVK_HD bool replay(PSeq *ps, SeqInfo &s, int side, u32 mask, int depth, bool to_end) const {
        int cur = 0, ref_pos = 0, len = 0, m = 0, skip = 0, last_ok = 1;
        const bool want_truth = side == 0;
        const Work &W = w();
        if (PIECES) { ps->ls[0] = 0; ps->src[0] = 0; }
        for (int oi = 0; oi < depth; ++oi) {
            if (is_truth(oi) == want_truth && ((mask >> oi) & 1)) {
                const VarInfo v = W.var[oi];
                const int vpos = v.pos;
                if (ref_pos <= vpos) {                        // compatible (:189)
                    if (m >= TS_MAXALT) return false;
                    len += vpos - cur;
                    if (PIECES) { ps->ls[2 * m + 1] = (u16)len; ps->src[2 * m + 1] = (u16)(v.aoff + v.l0); }
                    len += v.l1;
                    cur = vpos + v.l0;
                    if (PIECES) { ps->ls[2 * m + 2] = (u16)len; ps->src[2 * m + 2] = (u16)cur; }
                    m += 1;
                    ref_pos = cur;
                } else {
                    skip += v.alted;                          // edit_distance(allele0, allele1) (:199)
                    if (oi == depth - 1) last_ok = 0;
                }
            }
            const int sy = sync_pos(oi);
            if (ref_pos < sy) ref_pos = sy;
        }
        if (to_end && ref_pos < c.end - c.start) ref_pos = c.end - c.start;
        len += ref_pos - cur;
        if (PIECES) ps->ls[2 * m + 1] = (u16)len;
        s.len = len; s.ref_pos = ref_pos; s.skip = skip; s.n_alt = m; s.last_ok = last_ok;
        return true;
    }

    AVK_HD const u8 *piece_ptr(const PSeq &s, int k, int x) const { return ((k & 1) ? c.alle : c.ref) + s.src[k] + (x - s.ls[k]); }

    // equal leading bytes of A[x..] and B[y..]
    AVK_HD int lcp(const PSeq &A, int la, int x, const PSeq &B, int lb, int y) const {
        const int maxn = min_i(la - x, lb - y);
        if (maxn <= 0) return 0;
        int ka = 0, kb = 0;
        while (A.ls[ka + 1] <= x) ++ka;
        while (B.ls[kb + 1] <= y) ++kb;
        int total = 0;
        for (;;) {
            const int n = min_i(min_i(A.ls[ka + 1] - x, B.ls[kb + 1] - y), maxn - total);
            const u8 *pa = piece_ptr(A, ka, x), *pb = piece_ptr(B, kb, y);
            if (pa != pb) {                                   // same reference bytes otherwise: equal by construction
                int j = 0;
                while (j < n && pa[j] == pb[j]) ++j;
                if (j < n) return total + j;
            }
            total += n; x += n; y += n;
            if (total >= maxn) return total;
            while (A.ls[ka + 1] <= x) ++ka;
            while (B.ls[kb + 1] <= y) ++kb;
        }
    }

    // DWFALite::update (to_full == false, dynamic_wfa.rs:68-84) / finalize (:183-198) on wavefront wf with distance *ed.
    // returns false when the distance would exceed TS_EDCAP (the cluster is rejected)
    AVK_HD bool dwfa_run(u16 *wf, int *ed_io, const PSeq &A, int la, const PSeq &B, int lb, bool to_full) {
        int ed = *ed_io;
        for (;;) {
            int mb = -1, mo = -1, matched = 0;
            bool full = false;
            const int n = 2 * ed + 1;
            for (int i = 0; i < n; ++i) {                     // extend(): :94-130
                int d = wf[i];
                int boff = d + ed - i;
                if (boff < la && d < lb) {
                    const int ext = lcp(A, la, boff, B, lb, d);
                    d += ext; boff += ext; matched += ext;
                    wf[i] = (u16)d;
                }
                mb = max_i(mb, boff); mo = max_i(mo, d);
                full = full || (boff >= la && d >= lb);
            }
            ctr->cells += (u64)n; ctr->matched += (u64)matched;
            if (to_full ? full : (mb >= la || mo >= lb)) break;
            if (ed + 1 > TS_EDCAP) return false;
            // increase_edit_distance(): :152-168, in place from the top
            for (int i = n + 1; i >= 0; --i) {
                int v = 0;
                if (i < n) v = wf[i];
                if (i >= 1 && i - 1 < n) v = max_i(v, wf[i - 1] + 1);
                if (i >= 2 && i - 2 < n) v = max_i(v, wf[i - 2] + 1);
                wf[i] = (u16)v;
            }
            ed += 1;
        }
        *ed_io = ed;
        return true;
    }

    // ================================================================== begin: load the cluster
    // leaves phase = PH_SEARCH with the workspace loaded, or PH_FINISH with rc = an error status / TS_REJECT
    AVK_HD void begin(const u8 *digest, const u8 *contig, int start, int end, int mbf) {
        const int *hdr = (const int *)digest;
        phase = PH_FINISH;
        rc = hdr[PH_STATUS / 4];
        if (rc) return;
        rc = TS_REJECT;
        const int n = hdr[PH_N / 4];
        if (n > TS_MAXN || n < 1 || mbf > 255) return;
        // logical offsets are 16-bit, allele lengths 8-bit
        if ((end - start) + hdr[PH_SUM_L1 / 4] > 60000 || hdr[PH_SUM_ALLE / 4] > 60000) return;
        const int ns = hdr[PH_NSLOTS / 4];
        if (ns > TS_MAXSLOT) return;
        c.ref = contig + start; c.recs = digest + PH_SIZE; c.alle = c.recs + (size_t)VI_SIZE * n;
        c.start = start; c.end = end; c.N = n; c.nT = hdr[PH_N0 / 4]; c.nQ = hdr[PH_N1 / 4]; c.mbf = mbf; c.n_slots = ns;
        c.slot_types = 0;
        for (int k = 0; k < ns; ++k) c.slot_types |= (u32)digest[PH_SLOT_TYPE + k] << (4 * k);
        Work &W = w();
        u16 tm = 0;
        for (int oi = 0; oi < n; ++oi) {
            const u32 *r = (const u32 *)(c.recs + (size_t)VI_SIZE * oi);
            if (r[VI_L0 / 4] > 255u || r[VI_L1 / 4] > 255u) return;
            VarInfo v;
            v.pos = (u16)(r[VI_POS / 4] - (u32)start); v.aoff = (u16)r[VI_AOFF / 4];
            v.l0 = (u8)r[VI_L0 / 4]; v.l1 = (u8)r[VI_L1 / 4]; v.alted = (u8)r[VI_ALTED / 4]; v.pad = 0;   // alt_ed <= max(l0, l1)
            W.var[oi] = v;
            const u32 f = r[VI_FLAGS / 4];
            W.vtype[oi] = (u8)(f & 0xff); W.zyg[oi] = (u8)((f >> 8) & 0xff); W.slot[oi] = (u8)(f >> 24);
            if (f & 0x10000u) tm |= (u16)(1u << oi);
        }
        W.truth_mask = tm;
        // optimize_sequences: root (query_optimizer.rs:184-192)
        for (int i = 0; i <= n; ++i) W.bucket[i] = 0;
        nres = 0; best = 0xffffffffu; next_id = 1; qn = 0;
        { QEnt e; e.key = 0; e.a1 = 0; e.a2 = 0; e.depth = 0; e.ed1 = 0; e.ed2 = 0; e.pad = 0; W.q[qn++] = e; }
        rc = AVK_ST_OK;
        phase = PH_SEARCH;
    }
    AVK_HD void fail(int status) { rc = status; phase = PH_FINISH; }

    // ================================================================== one pop of optimize_sequences (query_optimizer.rs:203-328)
    AVK_HD void search_step() {
        Work &W = w();
        if (qn == 0) {                                                          // queue drained
            if (nres == 0) { fail(AVK_ST_NO_RESULT); return; }                  // :331
            score_begin();
            return;
        }
        const int n = c.N;
        int bi = 0;
        u32 bk = W.q[0].key;
        for (int i = 1; i < qn; ++i) { const u32 k = W.q[i].key; if (k < bk) { bk = k; bi = i; } }
        const QEnt e = W.q[bi];
        W.q[bi] = W.q[--qn];
        ctr->spops += 1;
        const u32 cost = e.key >> 16;
        if (cost > best) return;                                                // :204 strict
        const int oi = e.depth;
        if (W.bucket[oi] >= c.mbf) return;                                      // :222
        W.bucket[oi] += 1;
        const u32 pm[2] = {e.a1, e.a2};
        const int ped[2] = {e.ed1, e.ed2};
        PSeq &T = W.seq[0], &Q = W.seq[1];
        SeqInfo ti, qi;
        // parent wavefronts: one diagonal at the end of the shorter sequence when the distance is 0, else recomputed
        for (int h2 = 0; h2 < 2; ++h2) {
            if (ped[h2] == 0) {
                replay<false>(nullptr, ti, 0, pm[h2], oi, false); replay<false>(nullptr, qi, 1, pm[h2], oi, false);
                W.wf[h2][0] = (u16)min_i(ti.len, qi.len);
            } else {
                if (!replay<true>(&T, ti, 0, pm[h2], oi, false) || !replay<true>(&Q, qi, 1, pm[h2], oi, false)) { fail(TS_REJECT); return; }
                W.wf[h2][0] = 0;
                int ed = 0;
                if (!dwfa_run(W.wf[h2], &ed, T, ti.len, Q, qi.len, false) || ed != ped[h2]) { fail(TS_REJECT); return; }
            }
        }
        if (oi == n) {                                                          // :227-247 finalize_dwfa (haplotype_dwfa.rs:84-95)
            int fed[2], tsk[2], qsk[2];
            for (int h2 = 0; h2 < 2; ++h2) {
                if (!replay<true>(&T, ti, 0, pm[h2], n, true) || !replay<true>(&Q, qi, 1, pm[h2], n, true)) { fail(TS_REJECT); return; }
                int ed = ped[h2];
                if (!dwfa_run(W.wf[h2], &ed, T, ti.len, Q, qi.len, false) || !dwfa_run(W.wf[h2], &ed, T, ti.len, Q, qi.len, true)) { fail(TS_REJECT); return; }
                ctr->alignments += 1;
                fed[h2] = ed; tsk[h2] = ti.skip; qsk[h2] = qi.skip;
            }
            const u32 cc = (u32)(fed[0] + fed[1] + tsk[0] + tsk[1] + qsk[0] + qsk[1]);
            if (cc < best) { best = cc; nres = 0; }
            if (cc == best) {
                if (nres >= TS_RESCAP || (tsk[0] | tsk[1] | qsk[0] | qsk[1]) > 255) { fail(TS_REJECT); return; }
                ResEnt r;
                r.a1 = (u16)pm[0]; r.a2 = (u16)pm[1]; r.ed1 = (u8)fed[0]; r.ed2 = (u8)fed[1];
                r.tvs1 = (u8)tsk[0]; r.tvs2 = (u8)tsk[1]; r.qvs1 = (u8)qsk[0]; r.qvs2 = (u8)qsk[1]; r.p0 = r.p1 = 0;
                W.res[nres++] = r;
            }
            return;
        }
        const int z = W.zyg[oi];
        const bool tr = is_truth(oi);
        const bool het = (z == AVK_ZYG_UNPHASED_HET || z == AVK_ZYG_PHASED_HET01 || z == AVK_ZYG_PHASED_HET10);
        if (!het && z != AVK_ZYG_HOM_ALT) { fail(AVK_ST_BAD_ZYGOSITY); return; }   // assert_eq! :315
        const bool two = het && (!tr || z == AVK_ZYG_UNPHASED_HET);                // :269 both orientations, new ids
        for (int k = two ? 0 : 1; k < 2; ++k) {
            bool a1, a2;
            if (two) { a1 = k == 1; a2 = k == 0; }                              // (REF, ALT) first, then (ALT, REF)
            else if (het) { a1 = (z == AVK_ZYG_PHASED_HET10); a2 = !a1; }       // phased truth het :294-312
            else { a1 = true; a2 = true; }                                      // hom-alt :313-327
            const u32 cm[2] = {pm[0] | ((a1 ? 1u : 0u) << oi), pm[1] | ((a2 ? 1u : 0u) << oi)};
            int ced[2];
            u32 ccost = 0;
            for (int h2 = 0; h2 < 2; ++h2) {
                if (!replay<true>(&T, ti, 0, cm[h2], oi + 1, false) || !replay<true>(&Q, qi, 1, cm[h2], oi + 1, false)) { fail(TS_REJECT); return; }
                const int pn = 2 * ped[h2] + 1;
                for (int i = 0; i < pn; ++i) W.wf[2][i] = W.wf[h2][i];
                int ed = ped[h2];
                if (!dwfa_run(W.wf[2], &ed, T, ti.len, Q, qi.len, false)) { fail(TS_REJECT); return; }
                ced[h2] = ed;
                ccost += (u32)(ed + ti.skip + qi.skip);
            }
            u32 id;
            if (two) id = next_id++;
            else id = e.key & 0xffffu;
            if (ccost > 0xfffeu || id > 0xfffeu) { fail(TS_REJECT); return; }
            if (qn >= TS_QCAP) {                                                // garbage collection: entries the search would discard when popped (cost > best)
                int wq = 0;
                for (int i = 0; i < qn; ++i) if ((W.q[i].key >> 16) <= best) W.q[wq++] = W.q[i];
                qn = wq;
                if (qn >= TS_QCAP) { fail(TS_REJECT); return; }
            }
            QEnt ne;
            ne.key = (ccost << 16) | id; ne.a1 = (u16)cm[0]; ne.a2 = (u16)cm[1]; ne.depth = (u8)(oi + 1);
            ne.ed1 = (u8)ced[0]; ne.ed2 = (u8)ced[1]; ne.pad = 0;
            W.q[qn++] = ne;
        }
    }

    // ================================================================== scoring driver (waffle_solver.rs:169-265)
    // exact-GT scoring of every equal-best solution; first minimum wins, with the pruning argued in
    // RegionSolver::compare_score: a haplotype with ED 0 and nothing skipped scores 0 flips without a search; a solution whose
    // lower bound reaches the current minimum is not searched; a search stops at its budget; the first solution with both
    // haplotypes at 0 is the answer.
    AVK_HD static bool hap_zero(const ResEnt &r, int h2) { return h2 ? (r.ed2 + r.tvs2 + r.qvs2 == 0) : (r.ed1 + r.tvs1 + r.qvs1 == 0); }
    AVK_HD void score_begin() {
        const Work &W = w();
        best_r = 0; best_total = 0x7fffffff; best_keep[0] = best_keep[1] = 0;
        ri = 0; hi = nres;
        for (int i = 0; i < nres; ++i) if (hap_zero(W.res[i], 0) && hap_zero(W.res[i], 1)) { ri = i; hi = i + 1; break; }
        h = 0;
        score_next();
    }
    // advance (ri, h) until an exact search has to run (phase = PH_EXACT) or every solution is scored (phase = PH_FINISH)
    AVK_HD void score_next() {
        const Work &W = w();
        for (;;) {
            if (ri >= hi) { phase = PH_FINISH; rc = AVK_ST_OK; return; }
            const ResEnt r = W.res[ri];
            const bool z0 = hap_zero(r, 0), z1 = hap_zero(r, 1);
            if (h == 0) {
                if ((z0 ? 0 : 1) + (z1 ? 0 : 1) >= best_total) { ri += 1; continue; }
                total = 0; lost = false; keep[0] = keep[1] = 0;
            }
            if (h < 2 && !lost) {
                const u32 ha = h ? r.a2 : r.a1;
                if (h ? z1 : z0) { keep[h] = ha; h += 1; continue; }
                budget = best_total - total - ((h == 0 && !z1) ? 1 : 0);
                exact_begin(ha);
                return;
            }
            if (!lost && total < best_total) { best_total = total; best_r = ri; best_keep[0] = keep[0]; best_keep[1] = keep[1]; }
            ri += 1; h = 0;
        }
    }
    AVK_HD void exact_done(int errs) {
        lost = errs >= budget;
        total += errs;
        keep[h] = x_keep;
        h += 1;
        score_next();
    }

    // ================================================================== optimize_gt_alleles (exact_gt_optimizer.rs:108-357)
    // ha: bit oi set <=> the haplotype's input allele of order entry oi is ALT.  Result: x_keep (bit set <=> ALT kept).
    AVK_HD void exact_begin(u32 ha) {
        Work &W = w();
        hap_alt = ha; x_keep = 0;
        x_next_id = 1; best_err = 0x7fffffff; have_best = false;
        min_sync = 0; af_index = 0; af_counts = 0;
        xn = 0;
        { XEnt e; e.key = (31u << 22); e.keep = 0; e.depth = 0; e.pad = 0; W.x[xn++] = e; }
        phase = PH_EXACT;
    }
    AVK_HD void exact_step() {
        Work &W = w();
        const int n = c.N;
        if (xn == 0) {
            if (!have_best) { fail(AVK_ST_NO_RESULT); return; }                  // :345-348
            exact_done(best_err);
            return;
        }
        int bi = 0;
        u32 bk = W.x[0].key;
        for (int i = 1; i < xn; ++i) { const u32 k = W.x[i].key; if (k < bk) { bk = k; bi = i; } }
        const XEnt e = W.x[bi];
        W.x[bi] = W.x[--xn];
        ctr->xpops += 1;
        const int errors = (int)(e.key >> 27);
        const u32 eid = e.key & 0x3fffffu;
        if (errors >= budget && !have_best) { exact_done(budget); return; }      // nodes pop in non-decreasing error order
        if (errors >= best_err) return;                                          // :169 non-strict
        const int oi = e.depth;
        PSeq &T = W.seq[0], &Q = W.seq[1];
        SeqInfo ti, qi;
        if (oi == n) {                                                           // :180-192: finalize; exact <=> sequences equal
            if (!replay<true>(&T, ti, 0, e.keep, n, true) || !replay<true>(&Q, qi, 1, e.keep, n, true)) { fail(TS_REJECT); return; }
            ctr->cells += 1; ctr->alignments += 1;
            bool exact = ti.len == qi.len;
            if (exact) { const int m = lcp(T, ti.len, 0, Q, qi.len, 0); ctr->matched += (u64)m; exact = m == ti.len; }
            if (exact && errors < best_err) { best_err = errors; have_best = true; x_keep = e.keep; }
            return;
        }
        if (oi < min_sync) return;                                               // :194-197
        replay<false>(nullptr, ti, 0, e.keep, oi, false); replay<false>(nullptr, qi, 1, e.keep, oi, false);
        if (ti.len == qi.len && ti.ref_pos == qi.ref_pos) { min_sync = oi; af_counts = 0; af_index = oi; }   // is_synchronized :206-217 (alive => ed == 0)
        const int d0 = min_i(ti.len, qi.len);                                    // the single diagonal of an alive node
        const bool is_alt = (hap_alt >> oi) & 1;
        const bool do_alt = is_alt && !(oi < af_index);
        // REF allele: move, id kept (:257-273).  ALT allele: (REF, error) with id next_id, then (ALT, no error) with the
        // following id unless auto-failed (:274-306).
        u32 ids[2] = {eid, 0};
        if (is_alt) { ids[0] = x_next_id; if (do_alt) ids[1] = x_next_id + 1; x_next_id += do_alt ? 2 : 1; }
        for (int k = 0; k < (do_alt ? 2 : 1); ++k) {
            const bool alt = k == 1;
            const u32 kp = e.keep | ((alt ? 1u : 0u) << oi);
            const int cerr = errors + ((is_alt && !alt) ? 1 : 0);
            if (!replay<true>(&T, ti, 0, kp, oi + 1, false) || !replay<true>(&Q, qi, 1, kp, oi + 1, false)) { fail(TS_REJECT); return; }
            bool ok = true;
            if (alt) ok = (is_truth(oi) ? ti : qi).last_ok != 0;                 // incompatible ALT: success == false -> dropped
            if (ok) {                                                            // DWFA with max ED 0: extend the diagonal, an end must be reached
                const int m = lcp(T, ti.len, d0, Q, qi.len, d0);
                ctr->cells += 1; ctr->matched += (u64)m;
                ok = (d0 + m >= ti.len) || (d0 + m >= qi.len);
            }
            if (!ok) continue;
            const int good = (oi + 1) - cerr;
            if (ids[k] > 0x3ffffeu || cerr > 30) { fail(TS_REJECT); return; }
            if (xn >= TS_XCAP) {                                                 // garbage collection (RegionSolver::gc_queue)
                int wq = 0;
                for (int i = 0; i < xn; ++i) {
                    const XEnt g = W.x[i];
                    const bool dead = (int)(g.key >> 27) >= best_err || (g.depth != n && g.depth < min_sync);
                    if (!dead) W.x[wq++] = g;
                }
                xn = wq;
                if (xn >= TS_XCAP) { fail(TS_REJECT); return; }
            }
            XEnt ne;
            ne.key = ((u32)cerr << 27) | ((u32)(31 - good) << 22) | ids[k]; ne.keep = (u16)kp; ne.depth = (u8)(oi + 1); ne.pad = 0;
            W.x[xn++] = ne;
        }
        af_counts += 1;                                                          // :310-339
        if (af_counts >= 500) {
            if (af_index >= n) { fail(AVK_ST_NO_RESULT); return; }
            int wq = 0;
            for (int i = 0; i < xn; ++i) {
                const XEnt g = W.x[i];
                const bool set = g.depth > af_index;
                if (!set || !((g.keep >> af_index) & 1)) W.x[wq++] = g;
            }
            xn = wq;
            af_index += 1;
            af_counts = 0;
        }
    }

    // ================================================================== final scoring (waffle_solver.rs:226-522)
    // generate_allele_sequence (:726-778) of `side` for the haplotype whose ALT mask is `mask`; type_filter < 0 keeps all.
    // *failed: summed alt_ed of the overlapping (skipped) ALTs; *closed: ED(reference window, sequence) when known without
    // aligning (RegionSolver::build_hap_seq), else -1.  false: more than TS_MAXALT spliced ALTs.
    AVK_HD bool allele_seq(PSeq &ps, SeqInfo &s, int side, u32 mask, int type_filter, int *failed_out, int *closed_out) const {
        int cur = 0, len = 0, m = 0, failed = 0;
        int subm = 0, ins = 0, del = 0;
        bool open = false;
        const bool want_truth = side == 0;
        const Work &W = w();
        ps.ls[0] = 0; ps.src[0] = 0;
        for (int oi = 0; oi < c.N; ++oi) {
            if (is_truth(oi) != want_truth) continue;
            if (!((mask >> oi) & 1)) continue;                                   // REF allele: skipped entirely (:738-741)
            if (type_filter >= 0 && W.vtype[oi] != type_filter) continue;
            const VarInfo v = W.var[oi];
            const int vpos = v.pos;
            if (vpos < cur) { failed += v.alted; continue; }                     // :745-753
            if (m >= TS_MAXALT) return false;
            len += vpos - cur;
            ps.ls[2 * m + 1] = (u16)len; ps.src[2 * m + 1] = (u16)(v.aoff + v.l0);
            len += v.l1;
            cur = vpos + v.l0;
            ps.ls[2 * m + 2] = (u16)len; ps.src[2 * m + 2] = (u16)cur;
            m += 1;
            const bool anchored = c.alle[v.aoff + v.l0] == c.ref[vpos];
            if (v.l0 == 1 && v.l1 == 1) subm += anchored ? 0 : 1;
            else if (v.l0 == 1 && anchored) ins += v.l1 - 1;
            else if (v.l1 == 1 && anchored) del += v.l0 - 1;
            else open = true;
        }
        len += (c.end - c.start) - cur;
        ps.ls[2 * m + 1] = (u16)len;
        s.len = len; s.ref_pos = c.end - c.start; s.skip = failed; s.n_alt = m; s.last_ok = 1;
        int closed = -1;
        if (!open) {
            if (ins == 0 && del == 0) { if (subm <= 2) closed = subm; }
            else if (subm == 0 && (ins == 0 || del == 0)) closed = ins + del;
        }
        *failed_out = failed; *closed_out = closed;
        return true;
    }
    // global edit distance (wfa_ed, sequence_alignment.rs:9-13); -1: beyond TS_EDCAP
    AVK_HD int wfa_ed(const PSeq &A, int la, const PSeq &B, int lb) {
        Work &W = w();
        W.wf[2][0] = 0;
        int ed = 0;
        ctr->alignments += 1;
        if (!dwfa_run(W.wf[2], &ed, A, la, B, lb, true)) return -1;
        return ed;
    }
    // ED(reference window, S): closed form when known, else aligned against the window as a single piece (wf[0] is free by
    // now and holds its two-entry piece list)
    AVK_HD int ed_to_ref(const PSeq &S, int ls, int closed) {
        if (closed >= 0) { ctr->alignments += 1; ctr->cells += 1; return closed; }
        PSeq &R = *(PSeq *)w().wf[0];
        R.ls[0] = 0; R.ls[1] = (u16)(c.end - c.start); R.src[0] = 0;
        return wfa_ed(R, c.end - c.start, S, ls);
    }

    // fills `sol`; returns AVK_ST_OK, an error status, or TS_REJECT.  Nothing global is written here.
    AVK_HD int finish(Solution &sol) {
        Work &W = w();
        const int n = c.N;
        const ResEnt R = W.res[best_r];
        // ---- rows: joint + one per distinct variant type
        const int ns = c.n_slots;
        sol.n_rows = 1 + ns;
        sol.row_group[0] = 0;
        for (int k = 0; k < ns; ++k) sol.row_group[1 + k] = (u8)(1 + slot_type(k));
        for (int k = 0; k <= ns; ++k) for (int m = 0; m < AVK_N_METRICS; ++m) sol.rows[k][m] = 0;
        u32 slot_cnt[TS_MAXSLOT][2];
        u64 slot_tot[TS_MAXSLOT][2];
        for (int k = 0; k < TS_MAXSLOT; ++k) { slot_cnt[k][0] = slot_cnt[k][1] = 0; slot_tot[k][0] = slot_tot[k][1] = 0; }
        // ---- per-variant expected / observed (:226-258), GT / HAP / WEIGHTED_HAP (+ add_swap_benchmark :269), RECORD_BP totals
        sol.n = n;
        for (int oi = 0; oi < n; ++oi) {
            const bool tr = is_truth(oi);
            const int exp_ = (int)((R.a1 >> oi) & 1) + (int)((R.a2 >> oi) & 1);
            const int obs_ = (int)((best_keep[0] >> oi) & 1) + (int)((best_keep[1] >> oi) & 1);
            if (exp_ < obs_) return AVK_ST_TRUTH_FP;                              // assert! :322 (cannot happen: flips only remove ALTs)
            sol.exp[oi] = (u8)(tr ? exp_ : obs_);                                 // query entries are toggled (compare_benchmark.rs:109-123)
            sol.obs[oi] = (u8)(tr ? obs_ : exp_);
            const int side = tr ? 0 : 1, slot = W.slot[oi];
            const u64 wgt = W.var[oi].alted;
            gm_add(sol.rows[0], 2 * side, wgt, exp_, obs_);
            gm_add(sol.rows[1 + slot], 2 * side, wgt, exp_, obs_);
            slot_cnt[slot][side] += 1;
            slot_tot[slot][side] += (u64)(W.zyg[oi] == AVK_ZYG_HOM_ALT ? 2 : 1) * rec32(c, oi, VI_RAW);
        }
        // ---- basepair metrics (:335-449)
        PSeq &T = W.seq[0], &Q = W.seq[1], &F = W.seq[2];
        SeqInfo ti, qi, fi;
        for (int h2 = 0; h2 < 2; ++h2) {
            const u32 hm = h2 ? R.a2 : R.a1;
            int failT, failQ, closedT, closedQ;
            if (!allele_seq(T, ti, 0, hm, -1, &failT, &closedT) || !allele_seq(Q, qi, 1, hm, -1, &failQ, &closedQ)) return TS_REJECT;
            const int altT = ti.n_alt, altQ = qi.n_alt;
            // X = ED(ref, truth), Y = ED(ref, query), Z = ED(truth, query) = the optimizer's finalised distance
            const int xi = altT ? ed_to_ref(T, ti.len, closedT) : 0;
            if (xi < 0) return TS_REJECT;
            const u64 X = (u64)xi;
            const u64 Z = h2 ? R.ed2 : R.ed1;
            u64 Y;
            if (Z == 0) Y = X;
            else { const int yi = altQ ? ed_to_ref(Q, qi.len, closedQ) : 0; if (yi < 0) return TS_REJECT; Y = (u64)yi; }
            const u64 tp = X + Y - Z;
            u64 *bp = sol.rows[0] + AVK_M_BASEPAIR;
            bp[0] += tp; bp[1] += 2 * X - tp + 2 * (u64)failT; bp[2] += tp; bp[3] += 2 * Y - tp + 2 * (u64)failQ;
            for (int k = 0; k < ns; ++k) {
                const int ft = slot_type(k);
                if (!type_supported(ft)) continue;
                for (int side = 1; side >= 0; --side) {                           // query filter (:395-410) then truth filter (:422-437)
                    const int nf = (int)slot_cnt[k][side];
                    if (nf == 0) continue;
                    u64 f_tp, f_bad;
                    if (nf == (side ? c.nQ : c.nT)) {                             // filtered == full haplotype
                        f_tp = tp; f_bad = side ? (2 * Y - tp + 2 * (u64)failQ) : (2 * X - tp + 2 * (u64)failT);
                    } else {
                        int failF, closedF;
                        if (!allele_seq(F, fi, side, hm, ft, &failF, &closedF)) return TS_REJECT;
                        const u64 other_ref = side ? X : Y;                       // ED(ref, unfiltered other haplotype)
                        const bool alt_other = side ? (altT != 0) : (altQ != 0);
                        u64 Ef = 0, Zf = other_ref;
                        if (fi.n_alt) {
                            const int e1 = ed_to_ref(F, fi.len, closedF);
                            if (e1 < 0) return TS_REJECT;
                            Ef = (u64)e1;
                            if (alt_other) {
                                const int e2 = side ? wfa_ed(T, ti.len, F, fi.len) : wfa_ed(F, fi.len, Q, qi.len);
                                if (e2 < 0) return TS_REJECT;
                                Zf = (u64)e2;
                            } else Zf = Ef;
                        }
                        f_tp = other_ref + Ef - Zf;
                        f_bad = 2 * Ef - f_tp + 2 * (u64)failF;
                    }
                    u64 *g = sol.rows[1 + k] + AVK_M_BASEPAIR + 2 * side;
                    g[0] += f_tp; g[1] += f_bad;
                }
            }
        }
        // ---- add_record_basepair_stats (:455-522), wrapping u64 like a release build
        u32 mask = supported_type_mask();                                         // every supported type gets a (possibly all-zero) entry (:444)
        for (int k = 0; k < ns; ++k) mask |= 1u << slot_type(k);
        {
            u64 truth_total = 0, query_total = 0;
            for (int k = 0; k < ns; ++k) { truth_total += slot_tot[k][0]; query_total += slot_tot[k][1]; }
            u64 *bp = sol.rows[0] + AVK_M_BASEPAIR;
            const u64 tfn = bp[1], qfp = bp[3];
            const u64 ttp = 2 * truth_total - tfn, qtp = 2 * query_total - qfp;
            if (!(ttp >= bp[0]) || !(qtp >= bp[2])) return AVK_ST_TP_UNDERFLOW;
            u64 *rb = sol.rows[0] + AVK_M_RECORD_BP;
            rb[0] += ttp; rb[1] += tfn; rb[2] += qtp; rb[3] += qfp;
            for (int k = 0; k < ns; ++k) {
                u64 *g = sol.rows[1 + k];
                const u64 fn_ = g[AVK_M_BASEPAIR + 1], fp_ = g[AVK_M_BASEPAIR + 3];
                g[AVK_M_RECORD_BP + 0] += 2 * slot_tot[k][0] - fn_; g[AVK_M_RECORD_BP + 1] += fn_;
                g[AVK_M_RECORD_BP + 2] += 2 * slot_tot[k][1] - fp_; g[AVK_M_RECORD_BP + 3] += fp_;
            }
        }
        sol.ed1 = R.ed1; sol.ed2 = R.ed2;
        sol.type_mask = (u16)mask;
        return AVK_ST_OK;
    }
};

}  // namespace avk_ts
