// avk_spec_search.cuh -- optimize_sequences (query_optimizer.rs:166-365) for ONE dense cluster by ONE warp, up to 32 queue
// pops at a time.
//
// A dense cluster (10-20 variants, most of them heterozygous) is hundreds of queue pops, and a pop is a few thousand
// dependent scalar instructions: one warp -- or one thread -- popping them one after the other takes milliseconds while the
// rest of the GPU waits for it (the "dense tail" of a pass).  The pops are not as sequential as they look:
//
//   * priorities are (cost, node id), unique, and costs never decrease from a node to its children;
//   * a fork (both orientations of a heterozygous variant, query_optimizer.rs:269-292) gives its two children NEW ids,
//     larger than every id in the queue; a single child (phased truth het, hom-alt) keeps its parent's id;
//   * so while the minimum cost in the queue is c, the reference pops the cost-c entries in id order, and after each of them
//     immediately its single child, grandchild, ... for as long as their cost stays c (same id, still the minimum) -- a CHAIN
//     that ends in a fork, a finalised result, a child whose cost went up, or a node dropped by the branch quota.  Nothing
//     a chain produces is popped before the cost-c entries that were already queued (fork children have larger ids, all
//     other products cost more), and nothing it reads is written by an earlier chain except the per-depth quota counters.
//
// Hence a BATCH: the up to 32 cost-c entries with the smallest ids, one per lane, each lane running its chain on its own
// wavefront / sequence scratch, then a commit in lane (= id = reference) order: quota counters, results (first minimum and
// list order as in the sequential loop), new ids for fork children in chain order, queue pushes.  The quota is the one
// coupling: a batch is cut to the longest prefix of lanes for which no counter can reach max_branch_factor inside the batch
// (a single lane always runs: it checks the true counters itself), so no lane is ever dropped by a count another lane of
// the same batch added.  The search replays the reference's pops exactly -- same keys, ids, quota decisions, result order.
//
// Nodes are the 16-byte entries of the thread solver (avk_thread_solver.cuh): (cost, id) key, the two haplotypes' ALT masks,
// depth and edit distances; trackers are replayed from the masks, wavefronts of ED-0 parents are known in closed form and
// recomputed from scratch otherwise (DWFALite is path independent); sequences are never materialised.
//
// The equal-best results leave in the warp solver's format (RegionSolver::res_alle / res_num) and the existing team stage
// scores them.  Whatever does not fit (more than SP_MAXN variants, queue / result / edit-distance / piece capacity, an
// error status) is REJECTED: nothing has been written and the team stage solves the cluster from scratch as before.
//
// Plain scalar C++ apart from the warp driver at the end: the chain and commit code is compiled for the host by
// tests/sp_host.cpp and checked there against the CPU oracle's optimize_sequences, 32 simulated lanes at a time.
#pragma once
#include <cstddef>
#include "avk_layout.h"

namespace avk_sp {

#if defined(__CUDA_ARCH__)
// unaligned 32-bit little-endian read through a generic pointer
static __device__ __forceinline__ uint32_t ld4u_any(const uint8_t *p) {
    const uint32_t *w = (const uint32_t *)((uintptr_t)p & ~(uintptr_t)3);
    return __funnelshift_r(w[0], w[1], ((uint32_t)(uintptr_t)p & 3u) * 8u);
}
#endif

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
using namespace avk;

// capacities (overridable for the host harness' what-if runs)
#ifndef AVK_SP_QCAP
#define AVK_SP_QCAP 384
#endif
#ifndef AVK_SP_RESCAP
#define AVK_SP_RESCAP 64
#endif
#ifndef AVK_SP_EDCAP
#define AVK_SP_EDCAP 24
#endif
#ifndef AVK_SP_MAXALT
#define AVK_SP_MAXALT 12
#endif
#ifndef AVK_SP_XCAP
#define AVK_SP_XCAP 25
#endif
enum { SP_MAXN = 31, SP_QCAP = AVK_SP_QCAP, SP_RESCAP = AVK_SP_RESCAP, SP_EDCAP = AVK_SP_EDCAP, SP_MAXALT = AVK_SP_MAXALT, SP_MAXP = 2 * SP_MAXALT + 1, SP_WF = 2 * SP_EDCAP + 2, SP_LANES = 32 };
// dense blob: what the team stage reads back (slot = index in the dense list)
enum { SPB_NRES = 0, SPB_RES = 16, SPB_ALLE = 32, SPB_RES_STRIDE = SPB_ALLE + 24, SPB_SIZE = SPB_RES + SP_RESCAP * SPB_RES_STRIDE };
enum { SPB_NONE = -1 };   // n_res value: no search result here, solve from scratch
enum { SPB_SCORED = 4, SPB_KEEP0 = 8, SPB_KEEP1 = 12 };   // header: scored != 0 -> result 0 is THE solution and keep0 / keep1 its exact-GT alleles
enum { SP_XCAP = AVK_SP_XCAP, SP_MAXTASK = 2 * SP_RESCAP, SP_MAXSLOT = 4, SP_MAXMT = 40 };
enum { SPB_DONE = -2 };   // n_res value: the cluster is finished, every output has been written

struct VarInfo { u16 pos, aoff; u8 l0, l1, alted, zyg; };                      // pos relative to the region start
struct QEnt { u32 key; u32 a1, a2; u8 depth, ed1, ed2, pad; };                // key = cost << 16 | id
struct ResEnt { u32 a1, a2; u8 ed1, ed2, tvs1, tvs2, qvs1, qvs2, p0, p1; };
struct PSeq {
    u16 ls[SP_MAXP + 1];       // logical start of piece k; ls[n] == length
    u16 src[SP_MAXP];          // even k: reference position relative to the region start; odd k: offset into the allele bytes
};
struct Spec { u8 side, depth, to_end, pad; u32 mask; };
struct SeqInfo { int len, ref_pos, skip, n_alt, last_ok, plen, prp; };
enum { INIT_KEEP = 0, INIT_ZERO = 1, INIT_COPY = 2, INIT_CLOSED0 = 3, INIT_VALUE = 4 };
enum { CH_NONE = 0, CH_RESULT = 1, CH_PUSH1 = 2, CH_FORK = 3, CH_REJECT = 4 };

// what one chain hands to the commit
struct ChainOut {
    QEnt c0, c1;               // CH_PUSH1: c0 (id kept); CH_FORK: c0 then c1 (ids assigned by the commit)
    ResEnt r;                  // CH_RESULT
    u32 rcost;
    u8 kind, d_first, n_counted, pops;   // quota counters [d_first, d_first + n_counted) each get + 1; pops = entries popped
};

struct XEnt { u32 key; u32 keep; u8 depth, p0, p1, p2; };                      // optimize_gt_alleles: key = errors << 27 | (31 - good) << 22 | id
// what one exact-GT search hands to the scoring
struct XOut { u32 keep; u16 errs; u8 status, pad; };                           // status: 0 solved, 1 reject the cluster
struct MTask { u8 a_side, b_side, ok, pad; u32 a_mask, b_mask; u32 ed; };      // one alignment of the final metrics: ED(seq a, seq b), both replayed to the region end
// per-lane scratch
struct Scratch {
    union {
        u16 wf[3][SP_WF];      // search: parent hap 0 / hap 1, child
        XEnt xq[SP_XCAP];      // scoring: the lane's optimize_gt_alleles queue
    };
    PSeq seq[2];
    u16 pad[2];                // an odd number of 32-bit words: the 32 lanes' scratch areas start on 32 different banks
};
static_assert(sizeof(Scratch) % 4 == 0 && (sizeof(Scratch) / 4) % 2 == 1, "sizeof(Scratch) must be an odd number of words");

// per-cluster state shared by the warp
struct Shared {
    VarInfo var[SP_MAXN];
    u8 slot[SP_MAXN];               // metric-row slot of each order entry's variant type
    u32 type_bits[SP_MAXSLOT];      // bit oi: order entry oi has the variant type of slot k
    u32 truth_mask;
    u32 slot_types;                 // 4 bits per slot: variant type of metric-row slot k
    int N, nT, nQ, n_slots, wlen, mbf;
    int has_noop;                   // some record's ALT equals its REF (alt_ed == 0)
    int qn, nres;
    u32 best, next_id;
    u32 spops;
    int max_bucket;
    u16 bucket[SP_MAXN + 2];
    u16 cand[SP_QCAP];
    ResEnt res[SP_RESCAP];
    union {
        struct {               // search
            QEnt q[SP_QCAP];
            QEnt batch[SP_LANES];
            ChainOut out[SP_LANES];
        };
        struct {               // scoring of the equal-best results
            u8 task_r[SP_MAXTASK], task_h[SP_MAXTASK];
            XOut xout[SP_MAXTASK];
            int n_tasks, lo, hi;
            int best_r;
            u32 keep0, keep1, xpops, xcap;
            int n_mtasks;
            MTask mtask[SP_MAXMT];
            u32 bp[1 + SP_MAXSLOT][4];   // basepair counters: joint row, then one per slot
        };
    };
};
static_assert(sizeof(XEnt) * SP_XCAP <= sizeof(u16) * 3 * SP_WF, "the exact-GT queue must fit the wavefront scratch");

struct Counters { u64 cells, matched; u32 alignments; };

static AVK_HD inline int min_i(int a, int b) { return a < b ? a : b; }
static AVK_HD inline int max_i(int a, int b) { return a > b ? a : b; }
static AVK_HD inline int ctz32(u32 v) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
static AVK_HD inline Spec spec(int side, u32 mask, int depth, bool to_end) { Spec s; s.side = (u8)side; s.depth = (u8)depth; s.to_end = to_end ? 1 : 0; s.pad = 0; s.mask = mask; return s; }

// read-only view of one cluster for the chain code
struct View {
    const Shared *S;
    const u8 *ref;     // region window: base at absolute position start + x is ref[x]
    const u8 *alle;    // allele bytes of the digest
    const u8 *recs;    // the digest's VI_* records (merged order)
};

// HaplotypeTracker replay (haplotype_dwfa.rs:175-227): see Solver::replay_impl in avk_thread_solver.cuh -- an ALT of this
// side is compatible iff the last spliced ALT ends at or before it, so the walk visits the set bits of (mask & side) only.
static AVK_HD_NOINLINE bool replay(const Shared &S, PSeq *ps, SeqInfo *out, const Spec sp) {
    const int wlen = S.wlen, N = S.N;
    int cur = 0, len = 0, m = 0, skip = 0, last_ok = 1, plen = 0, prp = 0, ref_pos = 0;
    const bool pieces = ps != nullptr;
    if (pieces) { ps->ls[0] = 0; ps->src[0] = 0; }
    const int depth = sp.side < 2 ? (int)sp.depth : 0;
    if (depth > 0) {
        const u32 tmask = S.truth_mask;
        const u32 side_bits = sp.side == 0 ? tmask : ~tmask;
        const u32 last_bit = 1u << (depth - 1);
        u32 bits = sp.mask & side_bits & (last_bit | (last_bit - 1u));
        bool snap = false;
        while (bits) {
            const u32 low = bits & (0u - bits);
            bits ^= low;
            const int i = ctz32(low);
            const VarInfo v = S.var[i];
            const int vpos = v.pos;
            if (low == last_bit) {                            // state before the last replayed entry (the parent's)
                prp = depth == 1 ? 0 : (cur > vpos ? cur : vpos);
                plen = len + (prp - cur);
                snap = true;
            }
            if (cur <= vpos) {                                // compatible (haplotype_dwfa.rs:189)
                if (m >= SP_MAXALT) return false;
                len += vpos - cur;
                if (pieces) { ps->ls[2 * m + 1] = (u16)len; ps->src[2 * m + 1] = (u16)(v.aoff + v.l0); }
                len += v.l1;
                cur = vpos + v.l0;
                if (pieces) { ps->ls[2 * m + 2] = (u16)len; ps->src[2 * m + 2] = (u16)cur; }
                m += 1;
            } else {
                skip += v.alted;                              // edit_distance(allele0, allele1) (:199)
                if (low == last_bit) last_ok = 0;
            }
        }
        if (!snap) {
            const int lp = S.var[depth - 1].pos;
            prp = depth == 1 ? 0 : (cur > lp ? cur : lp);
            plen = len + (prp - cur);
        }
        const int sy = (depth == N) ? wlen : (int)S.var[depth].pos;   // sync point of the last entry (query_optimizer.rs:258-265)
        ref_pos = cur > sy ? cur : sy;
    }
    if ((sp.to_end || sp.side >= 2) && ref_pos < wlen) ref_pos = wlen;
    len += ref_pos - cur;
    if (pieces) ps->ls[2 * m + 1] = (u16)len;
    out->len = len; out->ref_pos = ref_pos; out->skip = skip; out->n_alt = m; out->last_ok = last_ok; out->plen = plen; out->prp = prp;
    return true;
}

static AVK_HD inline const u8 *piece_ptr(const View &V, const PSeq &s, int pk, int x) { return ((pk & 1) ? V.alle : V.ref) + s.src[pk] + (x - s.ls[pk]); }

// equal leading bytes of A[x..] and B[y..]
static AVK_HD_NOINLINE int lcp(const View &V, const PSeq &A, int la, int x, const PSeq &B, int lb, int y) {
    const int maxn = min_i(la - x, lb - y);
    if (maxn <= 0) return 0;
    int ka = 0, kb = 0;
    while (A.ls[ka + 1] <= x) ++ka;
    while (B.ls[kb + 1] <= y) ++kb;
    int total = 0;
    for (;;) {
        const int n = min_i(min_i(A.ls[ka + 1] - x, B.ls[kb + 1] - y), maxn - total);
        const u8 *pa = piece_ptr(V, A, ka, x), *pb = piece_ptr(V, B, kb, y);
        if (pa != pb) {                                   // same reference bytes otherwise: equal by construction
            int j = 0;
#if defined(__CUDA_ARCH__)
            while (j < n) {                                   // four bases per step (unaligned words: two aligned loads + funnel shift;
                const u32 x = ld4u_any(pa + j) ^ ld4u_any(pb + j);   //  every buffer read here has >= 8 bytes of slack behind it)
                if (x) { j += (__ffs((int)x) - 1) >> 3; break; }
                j += 4;
            }
            if (j > n) j = n;
#else
            while (j < n && pa[j] == pb[j]) ++j;
#endif
            if (j < n) return total + j;
        }
        total += n; x += n; y += n;
        if (total >= maxn) return total;
        while (A.ls[ka + 1] <= x) ++ka;
        while (B.ls[kb + 1] <= y) ++kb;
    }
}

// One alignment: build sequences a (baseline) and b (other), then DWFALite::update (dynamic_wfa.rs:68-84) if `update` and
// finalize (:183-198) if `finalize` on the wavefront `wf`, initialised per `init` (INIT_COPY: from `swf`; INIT_VALUE: one
// diagonal at v0).  Sequences are built in this lane's scratch; the wavefronts may live in another lane's.  false: capacity exceeded.
static AVK_HD_NOINLINE bool align(const View &V, Scratch &X, Counters &ctr, const Spec a, const Spec b, u16 *wf, int init, const u16 *swf, int ed_in, int v0,
                                  bool update, bool finalize, int *ed_out, SeqInfo *ia_out, SeqInfo *ib_out, int ed_cap = SP_EDCAP) {
    const Shared &S = *V.S;
    SeqInfo ia, ib;
    if (!replay(S, &X.seq[0], &ia, a) || !replay(S, &X.seq[1], &ib, b)) return false;
    *ia_out = ia; *ib_out = ib;
    const PSeq &A = X.seq[0], &B = X.seq[1];
    const int la = ia.len, lb = ib.len;
    int ed = ed_in;
    if (init == INIT_ZERO) { wf[0] = 0; ed = 0; }
    else if (init == INIT_CLOSED0) { wf[0] = (u16)min_i(ia.plen, ib.plen); ed = 0; }   // parent had ED 0: its one diagonal stood at the end of its shorter sequence
    else if (init == INIT_VALUE) { wf[0] = (u16)v0; ed = 0; }
    else if (init == INIT_COPY) { for (int i = 0; i < 2 * ed + 1; ++i) wf[i] = swf[i]; }
    bool fin_pass = !update;
    for (;;) {
        const int n = 2 * ed + 1;
        int mb = -1, mo = -1, matched = 0;
        bool full = false;
        for (int i = 0; i < n; ++i) {                          // extend() (:94-130)
            int d = wf[i];
            int boff = d + ed - i;
            if (boff < la && d < lb) {
                const int ext = lcp(V, A, la, boff, B, lb, d);
                d += ext; boff += ext; matched += ext;
                wf[i] = (u16)d;
            }
            mb = max_i(mb, boff); mo = max_i(mo, d);
            full = full || (boff >= la && d >= lb);
        }
        ctr.cells += (u64)n; ctr.matched += (u64)matched;
        const bool done = fin_pass ? full : (mb >= la || mo >= lb);
        if (done) {
            if (!fin_pass && finalize) { fin_pass = true; continue; }
            if (fin_pass) ctr.alignments += 1;
            *ed_out = ed;
            return true;
        }
        if (ed + 1 > ed_cap) { *ed_out = ed + 1; return false; }
        for (int i = n + 1; i >= 0; --i) {                     // increase_edit_distance() (:152-168): in place from the top
            int v = 0;
            if (i < n) v = wf[i];
            if (i >= 1 && i - 1 < n) v = max_i(v, wf[i - 1] + 1);
            if (i >= 2 && i - 2 < n) v = max_i(v, wf[i - 2] + 1);
            wf[i] = (u16)v;
        }
        ed += 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// A chain is run by a GROUP of G = 1, 2 or 4 adjacent lanes (G = 32 / batch size, rounded down): every lane of the group
// executes the chain's control flow on the same values, and where a pop needs several alignments that are independent of
// each other -- the two parent wavefronts, then the two haplotypes of each child -- they are dealt out over the group's
// lanes (job j to lane j mod G), each lane building its job's sequences in its own scratch, and the results are exchanged
// with shuffles.  The group's wavefront buffers are the 3 G buffers of its lanes' scratch areas.
struct Group {
    int G, sub;            // lanes in the group, this lane's index in it
    u32 mask;              // the group's lanes
    int base;              // first lane of the group
    Scratch *gs;           // scratch of the group's first lane (the others follow)
};
struct Job { Spec a, b; u16 *wf; const u16 *swf; int init, ed_in, v0; bool update, finalize; };
struct JobRes { int ed, askip, bskip; bool ok; };
static AVK_HD inline u16 *group_wf(const Group &g, int b) { return g.gs[b / 3].wf[b % 3]; }

static AVK_HD inline void run_jobs(const View &V, const Group &g, Scratch &X, Counters &ctr, const Job *jobs, int nj, JobRes *res) {
    for (int j = 0; j < nj; ++j) { res[j].ok = true; res[j].ed = 0; res[j].askip = 0; res[j].bskip = 0; }
    for (int j = g.sub; j < nj; j += g.G) {                                     // (one trip for every lane of the group: they stay converged)
        int ed = 0; SeqInfo ia, ib;
        ia.skip = ib.skip = 0;
        const Job &J = jobs[j];
        res[j].ok = align(V, X, ctr, J.a, J.b, J.wf, J.init, J.swf, J.ed_in, J.v0, J.update, J.finalize, &ed, &ia, &ib);
        res[j].ed = ed; res[j].askip = ia.skip; res[j].bskip = ib.skip;
    }
#if defined(__CUDA_ARCH__)
    if (g.G > 1) {
        __syncwarp(g.mask);                                                     // the wavefronts written above are read by other lanes' next jobs
        for (int j = 0; j < nj; ++j) {
            const int owner = g.base + j % g.G;
            res[j].ed = __shfl_sync(g.mask, res[j].ed, owner);
            res[j].askip = __shfl_sync(g.mask, res[j].askip, owner);
            res[j].bskip = __shfl_sync(g.mask, res[j].bskip, owner);
            res[j].ok = __shfl_sync(g.mask, (int)res[j].ok, owner) != 0;
        }
    }
#endif
}

// One chain: the popped entry `e` (cost c = the minimum in the queue) and, for as long as it has a single child of the same
// cost, that child.  Quota counters are read as they stood before the batch (see the file header).  Every lane of the group
// returns the same ChainOut.
static AVK_HD_NOINLINE void run_chain(const View &V, const Group &g, Scratch &X, Counters &ctr, QEnt e, ChainOut &o) {
    const Shared &S = *V.S;
    const int n = S.N;
    const u32 c = e.key >> 16;
    o.kind = CH_NONE; o.d_first = e.depth; o.n_counted = 0; o.pops = 0; o.rcost = 0;
    // wavefront buffers (indices into the group's pool): parent hap 0 / hap 1, children [k][h].  A lone lane has three
    // buffers: the children take the free one, and a single child's hap 1 the buffer hap 0's parent has just left.
    int pb0 = 0, pb1 = 1, cb[2][2] = {{2, 2}, {2, 2}};
    if (g.G > 1) { cb[0][0] = 2; cb[0][1] = 3; cb[1][0] = 4; cb[1][1] = 5; }
    bool have_wf = false;                                                       // the parent's wavefronts are in place (chain continuation)
    Job jobs[4];
    JobRes res[4];
    for (;;) {
        o.pops += 1;
        const int oi = e.depth;
        if ((int)S.bucket[oi] >= S.mbf) return;                                 // :222 dropped by the branch quota
        o.n_counted += 1;
        const u32 pm0 = e.a1, pm1 = e.a2;
        const int ped0 = e.ed1, ped1 = e.ed2;
        // wavefront of a parent haplotype with ED > 0: recomputed from scratch (path independence) unless this chain has just
        // computed it as its previous node's child
        if (!have_wf) {
            int nj = 0, exp_ed[2];
            for (int h = 0; h < 2; ++h) {
                const int pe = h ? ped1 : ped0;
                if (pe == 0) continue;
                const u32 pm = h ? pm1 : pm0;
                Job &J = jobs[nj];
                J.a = spec(0, pm, oi, false); J.b = spec(1, pm, oi, false); J.wf = group_wf(g, h ? pb1 : pb0); J.swf = nullptr;
                J.init = INIT_ZERO; J.ed_in = 0; J.v0 = 0; J.update = true; J.finalize = false;
                exp_ed[nj++] = pe;
            }
            if (nj) {
                run_jobs(V, g, X, ctr, jobs, nj, res);
                for (int j = 0; j < nj; ++j) if (!res[j].ok || res[j].ed != exp_ed[j]) { o.kind = CH_REJECT; return; }
            }
        }
        if (oi == n) {                                                          // :227-247 finalize_dwfa (haplotype_dwfa.rs:84-95)
            for (int h = 0; h < 2; ++h) {
                const u32 pm = h ? pm1 : pm0;
                const int pe = h ? ped1 : ped0;
                Job &J = jobs[h];
                J.a = spec(0, pm, n, true); J.b = spec(1, pm, n, true); J.wf = group_wf(g, h ? pb1 : pb0); J.swf = nullptr;
                J.init = INIT_KEEP; J.ed_in = pe; J.v0 = 0; J.update = true; J.finalize = true;
                if (pe == 0) {                                                  // closed-form parent diagonal: see PC_S_FINAL in avk_thread_solver.cuh
                    SeqInfo ti, qi;
                    if (!replay(S, nullptr, &ti, spec(0, pm, n, false)) || !replay(S, nullptr, &qi, spec(1, pm, n, false))) { o.kind = CH_REJECT; return; }
                    J.init = INIT_VALUE; J.v0 = min_i(ti.len, qi.len);
                }
            }
            run_jobs(V, g, X, ctr, jobs, 2, res);
            for (int h = 0; h < 2; ++h) if (!res[h].ok || res[h].askip > 255 || res[h].bskip > 255 || res[h].ed > 255) { o.kind = CH_REJECT; return; }
            o.rcost = (u32)(res[0].ed + res[1].ed + res[0].askip + res[1].askip + res[0].bskip + res[1].bskip);
            o.r.a1 = pm0; o.r.a2 = pm1; o.r.ed1 = (u8)res[0].ed; o.r.ed2 = (u8)res[1].ed;
            o.r.tvs1 = (u8)res[0].askip; o.r.tvs2 = (u8)res[1].askip; o.r.qvs1 = (u8)res[0].bskip; o.r.qvs2 = (u8)res[1].bskip; o.r.p0 = o.r.p1 = 0;
            o.kind = CH_RESULT;
            return;
        }
        const int z = S.var[oi].zyg;
        const bool tr = (S.truth_mask >> oi) & 1;
        const bool het = (z == AVK_ZYG_UNPHASED_HET || z == AVK_ZYG_PHASED_HET01 || z == AVK_ZYG_PHASED_HET10);
        if (!het && z != AVK_ZYG_HOM_ALT) { o.kind = CH_REJECT; return; }       // assert_eq! :315 -> the team stage reports it
        const bool two = het && (!tr || z == AVK_ZYG_UNPHASED_HET);             // :269 both orientations, new ids
        // children: HaplotypeDWFA::extend_variant (haplotype_dwfa.rs:46-67) on both haplotypes of each
        u32 cmask[2][2];
        int nj = 0;
        if (g.G == 1 && !two) { cb[1][0] = 3 - pb0 - pb1; cb[1][1] = pb0; }     // (three buffers: 0 + 1 + 2 == 3)
        if (g.G == 1 && two) { cb[0][0] = cb[0][1] = cb[1][0] = cb[1][1] = 3 - pb0 - pb1; }
        for (int k = two ? 0 : 1; k < 2; ++k) {
            bool a1, a2;
            if (two) { a1 = k == 1; a2 = k == 0; }                              // (REF, ALT) first, then (ALT, REF)
            else if (z != AVK_ZYG_HOM_ALT) { a1 = (z == AVK_ZYG_PHASED_HET10); a2 = !a1; }   // phased truth het :294-312
            else { a1 = true; a2 = true; }                                      // hom-alt :313-327
            cmask[k][0] = pm0 | ((a1 ? 1u : 0u) << oi); cmask[k][1] = pm1 | ((a2 ? 1u : 0u) << oi);
            for (int h = 0; h < 2; ++h) {
                const int pe = h ? ped1 : ped0;
                Job &J = jobs[nj++];
                J.a = spec(0, cmask[k][h], oi + 1, false); J.b = spec(1, cmask[k][h], oi + 1, false);
                J.wf = group_wf(g, cb[k][h]); J.swf = group_wf(g, h ? pb1 : pb0);
                J.init = pe == 0 ? INIT_CLOSED0 : INIT_COPY; J.ed_in = pe; J.v0 = 0; J.update = true; J.finalize = false;
            }
        }
        if (g.G == 1) {
            // a lone lane's jobs share buffers (hap 1 of a single child overwrites hap 0's parent wavefront, fork children reuse
            // the free buffer): they run one after the other in this order, which is what makes that safe
        }
        run_jobs(V, g, X, ctr, jobs, nj, res);
        QEnt child[2];
        for (int k = two ? 0 : 1, q = 0; k < 2; ++k, ++q) {
            const JobRes &r0 = res[2 * q], &r1 = res[2 * q + 1];
            if (!r0.ok || !r1.ok) { o.kind = CH_REJECT; return; }
            const int ccost = r0.ed + r0.askip + r0.bskip + r1.ed + r1.askip + r1.bskip;
            if (ccost > 0xfffe || r0.ed > 255 || r1.ed > 255) { o.kind = CH_REJECT; return; }
            QEnt ne;
            ne.key = ((u32)ccost << 16) | (two ? 0u : (e.key & 0xffffu));       // fork children: id assigned by the commit
            ne.a1 = cmask[k][0]; ne.a2 = cmask[k][1]; ne.depth = (u8)(oi + 1); ne.ed1 = (u8)r0.ed; ne.ed2 = (u8)r1.ed; ne.pad = 0;
            child[k] = ne;
        }
        if (two) { o.c0 = child[0]; o.c1 = child[1]; o.kind = CH_FORK; return; }
        if ((child[1].key >> 16) != c) { o.c0 = child[1]; o.kind = CH_PUSH1; return; }
        // same cost, same id: the reference pops it next; its wavefronts become the parent's
        { const int t0 = pb0, t1 = pb1; pb0 = cb[1][0]; pb1 = cb[1][1]; if (g.G > 1) { cb[1][0] = t0; cb[1][1] = t1; } }
        have_wf = true;
        e = child[1];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Commit of one batch in lane order: exactly what the sequential loop does between two pops (query_optimizer.rs:203-328).
// Queue garbage collection removes only entries the search would discard when popped (cost > best).
// returns false: capacity exceeded / a chain rejected -> the cluster is rejected.
static AVK_HD inline bool sp_push(Shared &S, const QEnt &ne) {
    if (S.qn >= SP_QCAP) {
        int wq = 0;
        for (int i = 0; i < S.qn; ++i) if ((S.q[i].key >> 16) <= S.best) S.q[wq++] = S.q[i];
        S.qn = wq;
        if (S.qn >= SP_QCAP) return false;
    }
    S.q[S.qn++] = ne;
    return true;
}
static AVK_HD_NOINLINE bool commit_batch(Shared &S, int nb) {
    for (int l = 0; l < nb; ++l) {
        const ChainOut &o = S.out[l];
        S.spops += o.pops;
        for (int d = o.d_first; d < o.d_first + o.n_counted; ++d) { S.bucket[d] += 1; S.max_bucket = max_i(S.max_bucket, (int)S.bucket[d]); }
        switch (o.kind) {
        case CH_NONE: break;
        case CH_RESULT:
            if (o.rcost < S.best) { S.best = o.rcost; S.nres = 0; }
            if (o.rcost == S.best) {
                if (S.nres >= SP_RESCAP) return false;
                S.res[S.nres++] = o.r;
            }
            break;
        case CH_PUSH1:
            if (!sp_push(S, o.c0)) return false;
            break;
        case CH_FORK: {
            if (S.next_id + 1 > 0xfffeu) return false;
            QEnt a = o.c0, b = o.c1;
            a.key |= S.next_id; b.key |= S.next_id + 1;
            S.next_id += 2;
            if (!sp_push(S, a) || !sp_push(S, b)) return false;
            break;
        }
        default: return false;
        }
    }
    return true;
}

// Loads the cluster from its digest (k_prep_fill).  false: outside the fast path's limits.
static AVK_HD_NOINLINE bool load_cluster(Shared &S, const u8 *digest, int start, int end, int mbf) {
    const int *hdr = (const int *)digest;
    if (hdr[PH_STATUS / 4]) return false;                                       // error status: the team stage writes it
    const int n = hdr[PH_N / 4];
    if (n > SP_MAXN || n < 1 || mbf <= 0 || mbf > 60000) return false;
    if ((end - start) + hdr[PH_SUM_L1 / 4] > 60000 || hdr[PH_SUM_ALLE / 4] > 60000) return false;   // logical offsets are 16-bit
    const u8 *recs = digest + PH_SIZE;
    u32 tm = 0;
    int noop = 0;
    for (int k = 0; k < SP_MAXSLOT; ++k) S.type_bits[k] = 0;
    for (int i = 0; i < n; ++i) {
        const u32 *r = (const u32 *)(recs + (size_t)VI_SIZE * i);
        if (r[VI_L0 / 4] > 255u || r[VI_L1 / 4] > 255u) return false;
        VarInfo v;
        v.pos = (u16)(r[VI_POS / 4] - (u32)start); v.aoff = (u16)r[VI_AOFF / 4];
        v.l0 = (u8)r[VI_L0 / 4]; v.l1 = (u8)r[VI_L1 / 4]; v.alted = (u8)r[VI_ALTED / 4];   // alt_ed <= max(l0, l1)
        if (v.alted == 0) noop = 1;
        const u32 f = r[VI_FLAGS / 4];
        v.zyg = (u8)((f >> 8) & 0xff);
        S.var[i] = v;
        if (f & 0x10000u) tm |= 1u << i;
        const u32 sl = f >> 24;
        S.slot[i] = (u8)sl;
        if (sl < (u32)SP_MAXSLOT) S.type_bits[sl] |= 1u << i;
    }
    S.truth_mask = tm; S.has_noop = noop;
    S.N = n; S.wlen = end - start; S.mbf = mbf;
    S.nT = hdr[PH_N0 / 4]; S.nQ = hdr[PH_N1 / 4]; S.n_slots = hdr[PH_NSLOTS / 4];
    S.slot_types = 0;
    for (int k = 0; k < S.n_slots && k < SP_MAXSLOT; ++k) S.slot_types |= (u32)digest[PH_SLOT_TYPE + k] << (4 * k);
    for (int i = 0; i <= n + 1; ++i) S.bucket[i] = 0;
    S.max_bucket = 0;
    S.nres = 0; S.best = 0xffffffffu; S.next_id = 1; S.spops = 0;
    QEnt root; root.key = 0; root.a1 = 0; root.a2 = 0; root.depth = 0; root.ed1 = 0; root.ed2 = 0; root.pad = 0;   // :184-192
    S.q[0] = root; S.qn = 1;
    return true;
}

// results -> dense blob (RegionSolver::res_alle / res_num layout: per order entry bit0 hap1 ALT, bit1 hap2 ALT; ed1 ed2 tvs1 tvs2 qvs1 qvs2)
static AVK_HD inline void store_result(u8 *blob, int ri, const ResEnt &r, int n) {
    u8 *p = blob + SPB_RES + (size_t)ri * SPB_RES_STRIDE;
    for (int i = 0; i < SP_MAXN; ++i) p[i] = i < n ? (u8)(((r.a1 >> i) & 1u) | (((r.a2 >> i) & 1u) << 1)) : 0;
    int *num = (int *)(p + SPB_ALLE);
    num[0] = r.ed1; num[1] = r.ed2; num[2] = r.tvs1; num[3] = r.tvs2; num[4] = r.qvs1; num[5] = r.qvs2;
}

// ---------------------------------------------------------------------------------------------------------------------
// Scoring of the equal-best results (waffle_solver.rs:169-265): optimize_gt_alleles (exact_gt_optimizer.rs:108-357) on both
// haplotypes of every result, first minimum of the summed errors wins.  The searches are independent of each other: one
// (result, haplotype) per lane, each an ordinary sequential best-first search in the lane's own queue (the code of the
// thread solver's PC_X_* states).  A haplotype with ED 0 and nothing skipped scores 0 flips without a search, and the first
// result with both haplotypes at 0 is the answer (RegionSolver::compare_score argues both).
// (in a cluster that holds a record whose ALT equals its REF the summed skip distance does not tell whether a variant was
// skipped: no shortcut there, see Solver::hap_zero in avk_thread_solver.cuh)
static AVK_HD inline bool hap_zero(const Shared &S, const ResEnt &r, int h) { return !S.has_noop && (h ? (r.ed2 + r.tvs2 + r.qvs2 == 0) : (r.ed1 + r.tvs1 + r.qvs1 == 0)); }

// longest common prefix from offset d0 of the two replayed sequences (DWFA with max ED 0, exact_gt_optimizer.rs:380)
static AVK_HD_NOINLINE bool prefix(const View &V, Scratch &X, Counters &ctr, const Spec a, const Spec b, int d0, int *m, SeqInfo *ia_out, SeqInfo *ib_out) {
    SeqInfo ia, ib;
    if (!replay(*V.S, &X.seq[0], &ia, a) || !replay(*V.S, &X.seq[1], &ib, b)) return false;
    *ia_out = ia; *ib_out = ib;
    *m = lcp(V, X.seq[0], ia.len, d0, X.seq[1], ib.len, d0);
    ctr.cells += 1; ctr.matched += (u64)*m;
    return true;
}

// optimize_gt_alleles on one haplotype: hap_alt = its ALT alleles in the result (by order entry).  `budget`: the search may
// stop as lost once the cheapest queued node has that many errors and nothing has been found.
static AVK_HD_NOINLINE void exact_gt_lane(const View &V, Scratch &X, Counters &ctr, u32 hap_alt, int budget, u32 xcap, u32 *xpops, XOut &o) {
    const Shared &S = *V.S;
    const int n = S.N;
    XEnt *xq = X.xq;
    int xn = 0, best_err = 0x7fffffff, min_sync = 0, af_index = 0, af_counts = 0;
    bool have_best = false;
    u32 next_id = 1, x_keep = 0, expansions = 0;
    o.status = 1; o.errs = 0; o.keep = 0; o.pad = 0;
    { XEnt e; e.key = (31u << 22); e.keep = 0; e.depth = 0; e.p0 = e.p1 = e.p2 = 0; xq[xn++] = e; }
    for (;;) {
        if (xn == 0) {
            if (!have_best) return;                                              // :345-348 no solution: the team stage reports it
            o.errs = (u16)best_err; o.keep = x_keep; o.status = 0;
            return;
        }
        int bi = 0;
        u32 bk = xq[0].key;
        for (int i = 1; i < xn; ++i) { const u32 kk = xq[i].key; if (kk < bk) { bk = kk; bi = i; } }
        const XEnt e = xq[bi];
        xq[bi] = xq[--xn];
        *xpops += 1;
        const int x_errors = (int)(e.key >> 27);
        if (x_errors >= budget && !have_best) { o.errs = (u16)budget; o.keep = 0; o.status = 0; return; }   // nodes pop in non-decreasing error order
        if (x_errors >= best_err) continue;                                      // :169 non-strict
        if (++expansions > xcap) return;                                         // stand-in for the 300 s bail (:174-176): the team stage reports it
        const int oi = e.depth;
        const u32 ekeep = e.keep;
        u32 id0 = e.key & 0x3fffffu;
        if (oi == n) {                                                           // :180-192: finalize; exact <=> sequences equal
            int m; SeqInfo ia, ib;
            if (!prefix(V, X, ctr, spec(0, ekeep, n, true), spec(1, ekeep, n, true), 0, &m, &ia, &ib)) return;
            ctr.alignments += 1;
            if (ia.len == ib.len && m == ia.len && x_errors < best_err) { best_err = x_errors; have_best = true; x_keep = ekeep; }
            continue;
        }
        if (oi < min_sync) continue;                                             // :194-197
        SeqInfo ti, qi;
        if (!replay(S, nullptr, &ti, spec(0, ekeep, oi, false)) || !replay(S, nullptr, &qi, spec(1, ekeep, oi, false))) return;
        if (ti.len == qi.len && ti.ref_pos == qi.ref_pos) { min_sync = oi; af_counts = 0; af_index = oi; }   // is_synchronized :206-217 (alive => ed == 0)
        const int d0 = min_i(ti.len, qi.len);                                    // the single diagonal of an alive node
        const bool is_alt = (hap_alt >> oi) & 1;
        const bool do_alt = is_alt && !(oi < af_index);
        // REF allele: move, id kept (:257-273).  ALT allele: (REF, error) with id next_id, then (ALT, no error) with the
        // following id unless auto-failed (:274-306).
        u32 id1 = 0;
        if (is_alt) { id0 = next_id; if (do_alt) id1 = next_id + 1; next_id += do_alt ? 2 : 1; }
        const bool tr = (S.truth_mask >> oi) & 1;
        for (int k = 0; k < (do_alt ? 2 : 1); ++k) {
            const bool alt = k == 1;
            const u32 kp = ekeep | ((alt ? 1u : 0u) << oi);
            int m; SeqInfo ia, ib;
            if (!prefix(V, X, ctr, spec(0, kp, oi + 1, false), spec(1, kp, oi + 1, false), d0, &m, &ia, &ib)) return;
            bool ok = true;
            if (alt) ok = (tr ? ia : ib).last_ok != 0;                           // incompatible ALT: success == false -> dropped
            if (ok) ok = (d0 + m >= ia.len) || (d0 + m >= ib.len);               // DWFA with max ED 0: an end must be reached
            if (!ok) continue;
            const int cerr = x_errors + ((is_alt && !alt) ? 1 : 0);
            const int good = (oi + 1) - cerr;
            const u32 id = alt ? id1 : id0;
            if (id > 0x3ffffeu || cerr > 30 || good > 31) return;
            if (xn >= SP_XCAP) {                                                 // garbage collection (RegionSolver::gc_queue)
                int wq = 0;
                for (int i = 0; i < xn; ++i) {
                    const XEnt g = xq[i];
                    const bool dead = (int)(g.key >> 27) >= best_err || (g.depth != n && g.depth < min_sync);
                    if (!dead) xq[wq++] = g;
                }
                xn = wq;
                if (xn >= SP_XCAP) return;
            }
            XEnt ne;
            ne.key = ((u32)cerr << 27) | ((u32)(31 - good) << 22) | id;
            ne.keep = kp; ne.depth = (u8)(oi + 1); ne.p0 = ne.p1 = ne.p2 = 0;
            xq[xn++] = ne;
        }
        af_counts += 1;                                                          // :310-339
        if (af_counts >= 500) {
            if (af_index >= n) return;
            int wq = 0;
            for (int i = 0; i < xn; ++i) {
                const XEnt g = xq[i];
                const bool set = g.depth > af_index;
                if (!set || !((g.keep >> af_index) & 1)) xq[wq++] = g;
            }
            xn = wq;
            af_index += 1;
            af_counts = 0;
        }
    }
}

// The searches to run: every non-zero haplotype of the results [lo, hi).  (Called after the search: its arrays alias.)
static AVK_HD_NOINLINE void score_prepare(Shared &S, u32 xcap) {
    int lo = 0, hi = S.nres;
    for (int i = 0; i < S.nres; ++i) if (hap_zero(S, S.res[i], 0) && hap_zero(S, S.res[i], 1)) { lo = i; hi = i + 1; break; }
    int nt = 0;
    for (int ri = lo; ri < hi; ++ri)
        for (int h = 0; h < 2; ++h)
            if (!hap_zero(S, S.res[ri], h)) { S.task_r[nt] = (u8)ri; S.task_h[nt] = (u8)h; nt += 1; }
    S.n_tasks = nt; S.lo = lo; S.hi = hi; S.xpops = 0; S.xcap = xcap;
}
// First minimum of the summed errors (waffle_solver.rs:264-265).  false: a search was rejected.
static AVK_HD_NOINLINE bool score_combine(Shared &S) {
    int best_total = 0x7fffffff, t = 0;
    S.best_r = S.lo; S.keep0 = S.keep1 = 0;
    for (int ri = S.lo; ri < S.hi; ++ri) {
        int total = 0;
        u32 keep[2];
        for (int h = 0; h < 2; ++h) {
            if (hap_zero(S, S.res[ri], h)) { keep[h] = h ? S.res[ri].a2 : S.res[ri].a1; continue; }
            const XOut &x = S.xout[t++];
            if (x.status) return false;
            total += x.errs; keep[h] = x.keep;
        }
        if (total < best_total) { best_total = total; S.best_r = ri; S.keep0 = keep[0]; S.keep1 = keep[1]; }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// Final metrics of the chosen solution (waffle_solver.rs:335-522), as in the thread solver's PC_F_* states and
// commit_solution: per haplotype X = ED(ref, truth), Y = ED(ref, query), Z = the optimizer's finalised distance; per
// supported type present the filtered sequence of one side against the full haplotype of the other.  The alignments these
// need are independent of each other: metrics_walk(collect) lists them (closed forms are resolved on the spot), the lanes
// run one each (metrics_align), metrics_walk(combine) -- the same walk -- consumes the distances in the same order.
static AVK_HD inline int popc32(u32 v) { int c = 0; while (v) { v &= v - 1; ++c; } return c; }
static AVK_HD inline int slot_type(const Shared &S, int k) { return (int)((S.slot_types >> (4 * k)) & 15u); }

// ED(reference window, haplotype of `side` with ALT mask `mask`) when it is known without aligning (RegionSolver::build_hap_seq,
// Solver::closed_form in avk_thread_solver.cuh); -1 otherwise
static AVK_HD_NOINLINE int closed_form(const View &V, int side, u32 mask) {
    const Shared &S = *V.S;
    int cur = 0, subm = 0, ins = 0, del = 0;
    bool open = false;
    u32 todo = mask & (side == 0 ? S.truth_mask : ~S.truth_mask) & (S.N >= 32 ? 0xffffffffu : ((1u << S.N) - 1u));
    while (todo) {
        const u32 low = todo & (0u - todo);
        todo ^= low;
        const VarInfo v = S.var[ctz32(low)];
        if ((int)v.pos < cur) continue;                                      // overlapping: skipped (:745-753)
        cur = v.pos + v.l0;
        const bool anchored = V.alle[v.aoff + v.l0] == V.ref[v.pos];
        if (v.l0 == 1 && v.l1 == 1) subm += anchored ? 0 : 1;
        else if (v.l0 == 1 && anchored) ins += v.l1 - 1;
        else if (v.l1 == 1 && anchored) del += v.l0 - 1;
        else open = true;
    }
    if (open) return -1;
    if (ins == 0 && del == 0) return subm <= 2 ? subm : -1;
    if (subm == 0 && (ins == 0 || del == 0)) return ins + del;
    return -1;
}

static AVK_HD_NOINLINE bool metrics_walk(const View &V, Shared &S, bool collect) {
    const int n = S.N;
    const ResEnt R = S.res[S.best_r];
    int t = 0;
    bool ok = true;
    // ED(a, b): queued while collecting, read back while combining
    auto need = [&](int a_side, u32 a_mask, int b_side, u32 b_mask) -> u32 {
        if (t >= SP_MAXMT) { ok = false; return 0; }
        MTask &m = S.mtask[t++];
        if (collect) { m.a_side = (u8)a_side; m.a_mask = a_mask; m.b_side = (u8)b_side; m.b_mask = b_mask; m.ok = 0; m.pad = 0; m.ed = 0; return 0; }
        if (!m.ok) ok = false;
        return m.ed;
    };
    if (!collect) for (int s2 = 0; s2 <= SP_MAXSLOT; ++s2) for (int m = 0; m < 4; ++m) S.bp[s2][m] = 0;
    for (int h = 0; h < 2 && ok; ++h) {
        const u32 f_mask = h ? R.a2 : R.a1;
        SeqInfo ti, qi;
        if (!replay(S, nullptr, &ti, spec(0, f_mask, n, true)) || !replay(S, nullptr, &qi, spec(1, f_mask, n, true))) return false;
        const int altT = ti.n_alt, altQ = qi.n_alt;
        const u32 failT = (u32)ti.skip, failQ = (u32)qi.skip;
        u32 X = 0, Y = 0;
        if (altT) { const int c = closed_form(V, 0, f_mask); X = c >= 0 ? (u32)c : need(2, 0, 0, f_mask); }
        const u32 Z = h ? R.ed2 : R.ed1;
        if (Z == 0) Y = X;
        else if (!altQ) Y = 0;
        else { const int c = closed_form(V, 1, f_mask); Y = c >= 0 ? (u32)c : need(2, 0, 1, f_mask); }
        const u32 tp = X + Y - Z;
        if (!collect) { S.bp[0][0] += tp; S.bp[0][1] += 2 * X - tp + 2 * failT; S.bp[0][2] += tp; S.bp[0][3] += 2 * Y - tp + 2 * failQ; }
        for (int k = 0; k < S.n_slots && ok; ++k) {
            if (!type_supported(slot_type(S, k))) continue;
            for (int side = 1; side >= 0; --side) {                             // query filter (:395-410) then truth filter (:422-437)
                const u32 side_bits = side ? ~S.truth_mask : S.truth_mask;
                const u32 fbits = S.type_bits[k] & side_bits;
                const int nf = popc32(fbits);
                if (nf == 0) continue;
                if (nf == (side ? S.nQ : S.nT)) {                               // filtered == full haplotype
                    if (!collect) { S.bp[1 + k][2 * side] += tp; S.bp[1 + k][2 * side + 1] += side ? (2 * Y - tp + 2 * failQ) : (2 * X - tp + 2 * failT); }
                    continue;
                }
                const u32 fm = f_mask & fbits;
                SeqInfo fi;
                if (!replay(S, nullptr, &fi, spec(side, fm, n, true))) return false;
                const u32 failF = (u32)fi.skip;
                const u32 other = side ? X : Y;                                 // ED(ref, unfiltered other haplotype)
                const int alt_other = side ? altT : altQ;
                u32 Ef = 0, Zf = other;
                if (fi.n_alt) {
                    const int c = closed_form(V, side, fm);
                    Ef = c >= 0 ? (u32)c : need(2, 0, side, fm);
                    if (!alt_other) Zf = Ef;
                    else Zf = side ? need(0, f_mask, 1, fm) : need(0, fm, 1, f_mask);   // ED(truth, F) resp. ED(F, query)
                }
                const u32 ftp = other + Ef - Zf;
                if (!collect) { S.bp[1 + k][2 * side] += ftp; S.bp[1 + k][2 * side + 1] += 2 * Ef - ftp + 2 * failF; }
            }
        }
    }
    if (collect) S.n_mtasks = t;
    return ok;
}
// one of the listed alignments: wfa_ed of the two replayed sequences (finalize from an empty wavefront)
// ed_cap < SP_EDCAP: a lane on its own gives up there (m.ok = 2) and the alignment is redone by the whole warp (metrics_align_warp)
static AVK_HD_NOINLINE void metrics_align(const View &V, Scratch &X, Counters &ctr, MTask &m, int ed_cap = SP_EDCAP) {
    int ed = 0; SeqInfo ia, ib;
    const int n = V.S->N;
    const bool ok = align(V, X, ctr, spec(m.a_side, m.a_mask, n, true), spec(m.b_side, m.b_mask, n, true), X.wf[0], INIT_ZERO, nullptr, 0, 0, false, true, &ed, &ia, &ib, ed_cap);
    m.ok = ok ? 1 : ((ed_cap < SP_EDCAP && ed > ed_cap) ? 2 : 0);
    m.ed = (u32)ed;
}

// Per-variant labels and metric rows of the finished cluster (commit_solution of avk_thread_solver.cuh):
//   sink.variant(order index, expected, observed), sink.metric(group, metric index, value)   (non-zero entries only)
// returns AVK_ST_OK, or the status the region ends with -- in which case the sink has not been called.
template <class Sink>
AVK_HD inline int commit_metrics(const View &V, const Shared &S, Sink &sink, u32 *ed1, u32 *ed2, u16 *type_mask) {
    const int n = S.N, ns = S.n_slots;
    const ResEnt R = S.res[S.best_r];
    u64 tot[SP_MAXSLOT][2];
    for (int s2 = 0; s2 < SP_MAXSLOT; ++s2) tot[s2][0] = tot[s2][1] = 0;
    u64 truth_total = 0, query_total = 0;
    for (int i = 0; i < n; ++i) {
        const int exp_ = (int)((R.a1 >> i) & 1) + (int)((R.a2 >> i) & 1);
        const int obs_ = (int)((S.keep0 >> i) & 1) + (int)((S.keep1 >> i) & 1);
        if (exp_ < obs_) return AVK_ST_TRUTH_FP;                                  // assert! :322 (cannot happen: flips only remove ALTs)
        const u64 v = (u64)(S.var[i].zyg == AVK_ZYG_HOM_ALT ? 2 : 1) * *(const u32 *)(V.recs + (size_t)VI_SIZE * i + VI_RAW);
        const int side = ((S.truth_mask >> i) & 1) ? 0 : 1;
        tot[S.slot[i]][side] += v;
        if (side == 0) truth_total += v; else query_total += v;
    }
    const u64 tfn = S.bp[0][1], qfp = S.bp[0][3];
    const u64 ttp = 2 * truth_total - tfn, qtp = 2 * query_total - qfp;
    if (!(ttp >= (u64)S.bp[0][0]) || !(qtp >= (u64)S.bp[0][2])) return AVK_ST_TP_UNDERFLOW;
    for (int i = 0; i < n; ++i) {                                                 // per-variant expected / observed (:226-258)
        const bool tr = (S.truth_mask >> i) & 1;
        const int exp_ = (int)((R.a1 >> i) & 1) + (int)((R.a2 >> i) & 1);
        const int obs_ = (int)((S.keep0 >> i) & 1) + (int)((S.keep1 >> i) & 1);
        sink.variant(i, tr ? exp_ : obs_, tr ? obs_ : exp_);
    }
    for (int s2 = -1; s2 < ns; ++s2) {                                            // rows: joint, then one per slot
        u64 gt[6] = {0, 0, 0, 0, 0, 0}, hap[4] = {0, 0, 0, 0}, whap[4] = {0, 0, 0, 0};
        for (int i = 0; i < n; ++i) {
            if (s2 >= 0 && S.slot[i] != s2) continue;
            const int col = ((S.truth_mask >> i) & 1) ? 0 : 2;
            const int exp_ = (int)((R.a1 >> i) & 1) + (int)((R.a2 >> i) & 1);
            const int obs_ = (int)((S.keep0 >> i) & 1) + (int)((S.keep1 >> i) & 1);
            const u64 wgt = S.var[i].alted;
            const int mn = exp_ < obs_ ? exp_ : obs_;
            hap[col] += (u64)mn; whap[col] += (u64)mn * wgt;
            if (exp_ == obs_) gt[col] += 1;
            else {
                hap[col + 1] += (u64)(exp_ - obs_); whap[col + 1] += (u64)(exp_ - obs_) * wgt; gt[col + 1] += 1;
                if (obs_ > 0) gt[col == 0 ? 4 : 5] += 1;
            }
        }
        const int g = s2 < 0 ? 0 : 1 + slot_type(S, s2);
        for (int m = 0; m < 6; ++m) if (gt[m]) sink.metric(g, AVK_M_GT + m, gt[m]);
        for (int m = 0; m < 4; ++m) {
            if (hap[m]) sink.metric(g, AVK_M_HAP + m, hap[m]);
            if (whap[m]) sink.metric(g, AVK_M_WEIGHTED_HAP + m, whap[m]);
            if (S.bp[1 + s2][m]) sink.metric(g, AVK_M_BASEPAIR + m, (u64)S.bp[1 + s2][m]);
        }
        u64 rb[4];
        if (s2 < 0) { rb[0] = ttp; rb[1] = tfn; rb[2] = qtp; rb[3] = qfp; }
        else {
            const u64 fn_ = S.bp[1 + s2][1], fp_ = S.bp[1 + s2][3];
            rb[0] = 2 * tot[s2][0] - fn_; rb[1] = fn_; rb[2] = 2 * tot[s2][1] - fp_; rb[3] = fp_;
        }
        for (int m = 0; m < 4; ++m) if (rb[m]) sink.metric(g, AVK_M_RECORD_BP + m, rb[m]);
    }
    u32 mask = supported_type_mask();                                             // every supported type gets a (possibly all-zero) entry (:444)
    for (int s2 = 0; s2 < ns; ++s2) mask |= 1u << slot_type(S, s2);
    *ed1 = R.ed1; *ed2 = R.ed2; *type_mask = (u16)mask;
    return AVK_ST_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Batch selection, sequential restatement (host harness; the warp version below does the same with ballots): the up to
// 32 minimum-cost entries with the smallest ids leave the queue into S.batch in id order, cut to the quota-safe prefix.
// returns the batch size; 0: the search is over (queue empty, or nothing left that costs <= best).
static AVK_HD inline int quota_safe_prefix(const Shared &S, int nb) {
    if (S.max_bucket + nb < S.mbf) return nb;
    for (int l = 1; l < nb; ++l)                              // lane 0 always runs
        for (int d = S.batch[l].depth; d <= S.N; ++d) {
            int before = 0;
            for (int k = 0; k < l; ++k) before += S.batch[k].depth <= d ? 1 : 0;
            if ((int)S.bucket[d] + before >= S.mbf) return l;
        }
    return nb;
}
#if !defined(__CUDA_ARCH__)
static inline int select_batch_host(Shared &S) {
    if (S.qn == 0) return 0;
    u32 cmin = 0xffffffffu;
    for (int i = 0; i < S.qn; ++i) cmin = S.q[i].key >> 16 < cmin ? S.q[i].key >> 16 : cmin;
    if (cmin > S.best) { S.spops += (u32)S.qn; S.qn = 0; return 0; }           // every remaining pop is discarded (:204)
    int nb = 0;
    for (;;) {                                                  // smallest keys first
        int bi = -1;
        for (int i = 0; i < S.qn; ++i) if ((S.q[i].key >> 16) == cmin && (bi < 0 || S.q[i].key < S.q[bi].key)) bi = i;
        if (bi < 0 || nb == SP_LANES) break;
        S.batch[nb++] = S.q[bi];
        S.q[bi] = S.q[--S.qn];
    }
    const int keep = quota_safe_prefix(S, nb);
    for (int l = keep; l < nb; ++l) S.q[S.qn++] = S.batch[l];   // back into the queue
    return keep;
}
#endif

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------------------------------
// Warp driver.  S: this warp's Shared; X: this lane's Scratch.  Returns true with S.res / S.nres filled, false = rejected.
__device__ __forceinline__ int select_batch_warp(Shared &S) {
    const int lane = threadIdx.x & 31;
    const int qn = S.qn;
    if (qn == 0) return 0;
    u32 cmin = 0xffffffffu;
    for (int i = lane; i < qn; i += 32) cmin = min(cmin, S.q[i].key >> 16);
    cmin = __reduce_min_sync(0xffffffffu, cmin);
    if (cmin > S.best) { __syncwarp(); if (lane == 0) { S.spops += (u32)qn; S.qn = 0; } __syncwarp(); return 0; }
    // candidates: queue indices of the cost-cmin entries
    int m = 0;
    for (int base = 0; base < qn; base += 32) {
        const int i = base + lane;
        const bool is = i < qn && (S.q[i].key >> 16) == cmin;
        const u32 bal = __ballot_sync(0xffffffffu, is);
        if (is) S.cand[m + __popc(bal & ((1u << lane) - 1u))] = (u16)i;
        m += __popc(bal);
    }
    __syncwarp();
    if (m > 32) {                                               // keep the 32 smallest ids: T = the 32nd smallest id
        u32 T = 0;
        for (int bit = 15; bit >= 0; --bit) {
            const u32 t2 = T | (1u << bit);
            int cnt = 0;
            for (int k = lane; k < m; k += 32) cnt += (S.q[S.cand[k]].key & 0xffffu) < t2 ? 1 : 0;
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (cnt < 32) T = t2;
        }
        int w = 0;
        for (int base = 0; base < m; base += 32) {              // in-place compaction (writes never pass the reads)
            const int k = base + lane;
            const u16 ci = k < m ? S.cand[k] : (u16)0;
            const bool is = k < m && (S.q[ci].key & 0xffffu) <= T;
            const u32 bal = __ballot_sync(0xffffffffu, is);
            __syncwarp();
            if (is) S.cand[w + __popc(bal & ((1u << lane) - 1u))] = ci;
            w += __popc(bal);
            __syncwarp();
        }
        m = w;                                                   // == 32
    }
    // rank by key among the (<= 32) candidates: lane l holds candidate l
    const bool have = lane < m;
    const int qi = have ? (int)S.cand[lane] : 0;
    const u32 key = have ? S.q[qi].key : 0xffffffffu;
    int rank = 0;
    for (int k = 0; k < 32; ++k) { const u32 ok = __shfl_sync(0xffffffffu, key, k); rank += ok < key ? 1 : 0; }
    if (have) S.batch[rank] = S.q[qi];
    __syncwarp();
    int nb = m;
    if (S.max_bucket + nb >= S.mbf) {                           // quota-safe prefix (quota_safe_prefix, one lane per batch entry)
        bool bad = false;
        const int myd = lane < nb ? (int)S.batch[lane].depth : 0x7fffffff;
        for (int d = 0; d <= S.N; ++d) {
            const u32 le = __ballot_sync(0xffffffffu, myd <= d);
            if (lane < nb && lane >= 1 && d >= myd && (int)S.bucket[d] + __popc(le & ((1u << lane) - 1u)) >= S.mbf) bad = true;
        }
        const u32 bb = __ballot_sync(0xffffffffu, bad);
        if (bb) nb = min(nb, __ffs(bb) - 1);
    }
    // remove the chosen entries from the queue: mark, then compact
    const bool chosen = have && rank < nb;
    if (chosen) S.q[qi].key = 0xffffffffu;
    __syncwarp();
    int w = 0;
    for (int base = 0; base < qn; base += 32) {
        const int i = base + lane;
        QEnt e;
        bool keep = false;
        if (i < qn) { e = S.q[i]; keep = e.key != 0xffffffffu; }
        const u32 bal = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) S.q[w + __popc(bal & ((1u << lane) - 1u))] = e;
        w += __popc(bal);
        __syncwarp();
    }
    if (lane == 0) S.qn = w;
    __syncwarp();
    return nb;
}

__device__ __noinline__ bool search_warp(Shared &S, Scratch &X, const View &V, Counters &ctr) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        const int nb = select_batch_warp(S);
        if (nb == 0) break;
        // a small batch leaves lanes free: groups of 2 or 4 lanes per chain share the alignments of each pop
        Group g;
        g.G = nb <= 8 ? 4 : (nb <= 16 ? 2 : 1);
        g.sub = lane % g.G; g.base = lane - g.sub; g.mask = (g.G == 32 ? 0xffffffffu : ((1u << g.G) - 1u)) << g.base; g.gs = &X - g.sub;
        const int chain = lane / g.G;
        if (chain < nb) {
            ChainOut o;
            run_chain(V, g, X, ctr, S.batch[chain], o);
            if (g.sub == 0) S.out[chain] = o;
        }
        __syncwarp();
        bool ok = true;
        if (lane == 0) ok = commit_batch(S, nb);
        ok = __shfl_sync(0xffffffffu, ok, 0);
        __syncwarp();
        if (!ok) return false;
    }
    return S.nres > 0;                                          // no result: the team stage reports AVK_ST_NO_RESULT
}
// scoring after search_warp: S.best_r / keep0 / keep1, or false = rejected
__device__ __noinline__ bool score_warp(Shared &S, Scratch &X, const View &V, Counters &ctr, u32 xcap) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (lane == 0) score_prepare(S, xcap);
    __syncwarp();
    const int nt = S.n_tasks;
    u32 xp = 0;
    for (int base = 0; base < nt; base += 32) {
        const int t = base + lane;
        if (t < nt) {
            const ResEnt &r = S.res[S.task_r[t]];
            exact_gt_lane(V, X, ctr, S.task_h[t] ? r.a2 : r.a1, 0x7fffffff, xcap, &xp, S.xout[t]);
        }
        __syncwarp();
    }
    xp = __reduce_add_sync(0xffffffffu, xp);
    bool ok = true;
    if (lane == 0) { S.xpops = xp; ok = score_combine(S); }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    __syncwarp();
    return ok;
}
// final metrics after score_warp: S.bp filled, or false = outside the fast path's limits
// wfa_ed of one listed alignment by the WHOLE warp: the lanes take the diagonals of a wavefront (a scalar lane needs (ED + 1)^2
// sequential longest-common-prefix walks; with ED around 20 that is half a millisecond).  X0 = lane 0's scratch: sequences in
// seq[0 / 1], wavefront ping-pong in wf[0 / 1].  Same recurrence and stop rule as align() with update == false, finalize == true.
enum { SP_EDCAP_LANE = 6 };
__device__ __noinline__ void metrics_align_warp(const View &V, Scratch &X0, Counters &ctr, MTask &m) {
    const int lane = threadIdx.x & 31;
    const Shared &S = *V.S;
    const int n = S.N;
    SeqInfo ia, ib;
    bool built = true;
    if (lane == 0) built = replay(S, &X0.seq[0], &ia, spec(m.a_side, m.a_mask, n, true)) && replay(S, &X0.seq[1], &ib, spec(m.b_side, m.b_mask, n, true));
    built = __shfl_sync(0xffffffffu, (int)built, 0) != 0;
    const int la = __shfl_sync(0xffffffffu, ia.len, 0), lb = __shfl_sync(0xffffffffu, ib.len, 0);
    __syncwarp();
    if (!built) { if (lane == 0) { m.ok = 0; m.ed = 0; } __syncwarp(); return; }
    const PSeq &A = X0.seq[0], &B = X0.seq[1];
    u16 *cur = X0.wf[0], *nxt = X0.wf[1];
    if (lane == 0) cur[0] = 0;
    __syncwarp();
    int ed = 0;
    bool ok = false;
    for (;;) {
        const int nd = 2 * ed + 1;
        bool full = false;
        int matched = 0;
        for (int i = lane; i < nd; i += 32) {                      // extend() (dynamic_wfa.rs:94-130), one diagonal per lane
            int d = cur[i];
            int boff = d + ed - i;
            if (boff < la && d < lb) { const int ext = lcp(V, A, la, boff, B, lb, d); d += ext; boff += ext; matched += ext; cur[i] = (u16)d; }
            full = full || (boff >= la && d >= lb);
        }
        ctr.matched += (u64)matched;
        if (lane == 0) ctr.cells += (u64)nd;
        __syncwarp();
        if (__any_sync(0xffffffffu, full)) { ok = true; break; }
        if (ed + 1 > SP_EDCAP) break;
        for (int i = lane; i < nd + 2; i += 32) {                  // increase_edit_distance() (:152-168) into the other buffer
            int v = 0;
            if (i < nd) v = cur[i];
            if (i >= 1 && i - 1 < nd) v = max_i(v, cur[i - 1] + 1);
            if (i >= 2 && i - 2 < nd) v = max_i(v, cur[i - 2] + 1);
            nxt[i] = (u16)v;
        }
        __syncwarp();
        u16 *t = cur; cur = nxt; nxt = t;
        ed += 1;
    }
    if (lane == 0) { m.ok = ok ? 1 : 0; m.ed = (u32)ed; if (ok) ctr.alignments += 1; }
    __syncwarp();
}

__device__ __noinline__ bool metrics_warp(Shared &S, Scratch &X, const View &V, Counters &ctr) {
    const int lane = threadIdx.x & 31;
    if (S.n_slots > SP_MAXSLOT) return false;
    bool ok = true;
    if (lane == 0) ok = metrics_walk(V, S, true);
    ok = __shfl_sync(0xffffffffu, ok, 0);
    __syncwarp();
    if (!ok) return false;
    const int nt = S.n_mtasks;
    for (int base = 0; base < nt; base += 32) {                    // one alignment per lane, up to a small edit distance ...
        if (base + lane < nt) metrics_align(V, X, ctr, S.mtask[base + lane], SP_EDCAP_LANE);
        __syncwarp();
    }
    for (int t = 0; t < nt; ++t) {                                 // ... and the few that go beyond it one after the other, 32 diagonals at a time
        if (S.mtask[t].ok != 2) continue;                          // (warp-uniform: shared memory)
        metrics_align_warp(V, (&X)[-lane], ctr, S.mtask[t]);
    }
    if (lane == 0) ok = metrics_walk(V, S, false);
    ok = __shfl_sync(0xffffffffu, ok, 0);
    __syncwarp();
    return ok;
}
#endif

}  // namespace avk_sp
