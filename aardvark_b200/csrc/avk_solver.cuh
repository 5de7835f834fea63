// avk_solver.cuh -- per-cluster solvers, one warp per cluster.
//
//   RegionSolver::optimize()       <- optimize_sequences      src/query_optimizer.rs:166-365
//   RegionSolver::exact_gt()       <- optimize_gt_alleles     src/exact_gt_optimizer.rs:108-357
//   RegionSolver::solve_compare()  <- solve_compare_region    src/waffle_solver.rs:122-284
//                                     compare_expected_observed :296-327, add_basepair_stats :335-449,
//                                     add_record_basepair_stats :455-522, generate_allele_sequence :726-778
//   RegionSolver::solve_merge()    <- solve_merge_region      src/merge_solver.rs:110-223
//
// The sequential best-first searches of the reference are replayed exactly: the same
// (unique) priority keys, the same node-id assignment, the same quotas and prunes, so the
// same equal-cost solution is reported.  What differs is the machine mapping: a warp keeps
// the whole problem of one cluster in a private workspace ("arena": shared memory in the
// common tiers, global memory for the rare oversized cluster):
//   * the reference window, staged once with a TMA bulk copy,
//   * the cluster's variants in merged processing order and their allele bytes,
//   * the search nodes (haplotype sequences + wavefronts), the key array of the queue,
//   * the metric rows being accumulated.
// Pops are a warp-wide scan + REDUX min over the unsorted key array, clones are warp-wide
// copies, scoring is the warp-cooperative DWFA of avk_device.cuh.  The solver object itself
// lives in shared memory (one per warp), never in a local-memory stack frame.
#pragma once
#include "avk_device.cuh"
#include "../../include/aardvark_b200.h"

namespace avk {

struct DevBatch {
    u64 n_regions;
    u32 n_inputs;
    const u64 *region_id;
    const u32 *contig, *start, *end;
    const u64 *var_off;
    const u32 *pos;
    const u8 *vtype, *zyg;
    const u32 *raw, *aoff, *l0, *l1;
    const u8 *pool;
    const u8 *const *contig_ptr;
    const u64 *contig_len;
    u32 n_contigs;
    const u32 *alt_ed;   // per variant, filled by k_alt_ed
};

struct DevCompareOut {
    int *status;
    u32 *ed1, *ed2;
    u64 *region_metrics;   // [n][13][22]
    uint16_t *type_mask;
    u8 *vexp, *vobs, *vcls;
    const u64 *seq_off;    // may be NULL
    u32 *seq_len;
    u8 *seq_pool;
};

struct DevMergeOut {
    int *status;
    u8 *cls, *n_idx, *idx;
};

enum { AL_UNSET = 0, AL_REF = 1, AL_ALT = 2 };
enum { SOLVE_OK = 0, SOLVE_WORKSPACE = -1 };   // internal; positive values are AVK_ST_*

__device__ __forceinline__ int align_up(int x, int a) { return (x + a - 1) / a * a; }

struct HapState {   // HaplotypeDWFA scalars (haplotype_dwfa.rs:17-24,145-154)
    int t_ref_pos, q_ref_pos, t_len, q_len, t_skip, q_skip, ed, pad;
};

// One variant of the cluster, in merged processing order (order_variants, query_optimizer.rs:372-381).
struct VInfo {
    u32 pos, l0, l1;
    u32 aoff;      // offset of allele0 in alle_base (allele1 follows it)
    u32 alt_ed;    // ED(allele0, allele1)  (variants.rs:413-415)
    u32 raw;       // raw_allele_space
    u32 gv;        // index in the batch variant table
    u8 type, zyg, is_truth, slot;   // slot = metric-row slot of this variant's type
};

static __device__ __forceinline__ bool type_supported(int t) {   // SUPPORTED_VARIANT_TYPES waffle_solver.rs:82-91
    return t == AVK_VT_SNV || t == AVK_VT_INSERTION || t == AVK_VT_DELETION || t == AVK_VT_INDEL ||
           t == AVK_VT_TR_CONTRACTION || t == AVK_VT_TR_EXPANSION || t == AVK_VT_SV_DELETION || t == AVK_VT_SV_INSERTION;
}

struct RegionSolver {
    // ---- problem (warp-uniform) ----
    const DevBatch *bp;
    WorkAcc work;
    const u8 *ref;        // ref[pos] addresses absolute contig positions (staged window or global contig)
    const u8 *alle_base;  // allele bytes (staged copy or the global pool)
    int start, end;       // region window
    int nv[2];
    int N;
    int mbf;
    bool stage;           // shared-memory tier: stage window + alleles into the arena
    // ---- arena ----
    u8 *arena;
    long long arena_bytes;
    VInfo *vinfo;         // [N]
    int *bucket;          // [N+1]
    u8 *res_alle;         // [res_cap][Npad] bit0 hap1 ALT, bit1 hap2 ALT (by order index)
    int *res_num;         // [res_cap][6] ed1 ed2 tvs1 tvs2 qvs1 qvs2
    u8 *hap_alle;         // [Npad] input alleles of exact_gt by order index
    u8 *cur_obs;          // [2][Npad]
    u8 *best_obs;         // [2][Npad]
    u64 *mrows;           // [1 + n_slots][22] metric rows: joint, then one per distinct variant type
    u32 *slot_cnt;        // [n_slots][2] number of truth / query variants of that type
    u64 *slot_tot;        // [n_slots][2] zygosity-weighted raw allele space of that type (RECORD_BP)
    u8 slot_type[AVK_N_VARIANT_TYPES];
    int n_slots;
    u8 *dyn;              // start of the re-partitionable part
    long long dyn_bytes;
    int Npad, seq_cap, wf_cap;
    int n_res, res_cap;
    int region_off;       // first arena byte after the mbarrier (+ staged window)
    u32 tma_phase;        // mbarrier parity of the next window load
    // queue + slots (partitioned per phase)
    u64 *qkeys;
    u32 *qslot;
    u32 *freel;
    u8 *nodes;
    int stride, max_slots, qn, nfree;

    // ------------------------------------------------------------------ helpers
    __device__ __forceinline__ int sync_pos(int oi) const {   // query_optimizer.rs:258-265
        return (oi == N - 1) ? end : (int)vinfo[oi + 1].pos;
    }

    // Start of a region: in shared-memory tiers the reference window is staged into the arena with a
    // TMA bulk copy (cp.async.bulk, 16-byte aligned superset of [start, end)), and `ref` is rebased so
    // that ref[pos] still addresses absolute contig positions.  Bytes [0,16) of the arena hold the
    // warp's mbarrier and are never touched by generic stores.
    __device__ __noinline__ bool begin_region(const u8 *contig) {
        region_off = 16;
        ref = contig;
        if (!stage) return true;
        const int a0 = start & ~15;
        const int bytes = align_up(end - a0 + 16, 16);        // >= 16 bytes of slack for ld4u
        if (16LL + bytes + 2048 > arena_bytes) return false;
        u8 *win = arena + 16;
        tma_window_load(win, contig + a0, (u32)bytes, (u64 *)arena, tma_phase);
        ref = win - a0;
        region_off = 16 + bytes;
        return true;
    }

    // per-lane partial validation of one variant list (same rules as the oracle's list_valid)
    __device__ __noinline__ bool validate_list(u64 v0, int n, int &sum_l1, int &b0, int &sum_alle) const {
        const DevBatch &b = *bp;
        bool invalid = false;
        #pragma unroll 1
        for (int i = lane_id(); i < n; i += 32) {
            const u64 gv = v0 + i;
            const u32 l0 = b.l0[gv], l1 = b.l1[gv], p = b.pos[gv];
            invalid = invalid || l0 == 0 || l1 == 0 || b.vtype[gv] >= AVK_N_VARIANT_TYPES || b.zyg[gv] > AVK_ZYG_HOM_ALT;
            invalid = invalid || (long long)p < start || (long long)p + l0 > (long long)end;
            if (i > 0) invalid = invalid || b.pos[gv - 1] > p;
            sum_l1 += (int)min(l1, 1u << 24);
            b0 += (int)min(max(l0, l1), 1u << 24);
            sum_alle += (int)min(l0, 1u << 24) + (int)min(l1, 1u << 24);
        }
        return invalid;
    }

    // Validate and load one (truth list, query list) pair of region r: lays out the arena, fills vinfo
    // in merged order, stages the allele bytes, assigns metric-row slots.
    __device__ __noinline__ int setup_pair(u64 r, u32 ki, u32 kj, bool want_metrics, bool &bad) {
        const DevBatch &b = *bp;
        const int lane = lane_id();
        const u32 K = b.n_inputs;
        const u64 v0[2] = {b.var_off[r * K + ki], b.var_off[r * K + kj]};
        nv[0] = (int)(b.var_off[r * K + ki + 1] - v0[0]);
        nv[1] = (int)(b.var_off[r * K + kj + 1] - v0[1]);
        N = nv[0] + nv[1];
        int sum_l1 = 0, b0 = 0, sum_alle = 0;
        bool invalid = false;
        #pragma unroll 1
        for (int side = 0; side < 2; ++side) invalid = validate_list(v0[side], nv[side], sum_l1, b0, sum_alle) || invalid;
        bad = __any_sync(AVK_FULL, invalid);
        if (bad) return SOLVE_OK;
        sum_l1 = __reduce_add_sync(AVK_FULL, sum_l1);
        b0 = __reduce_add_sync(AVK_FULL, b0);
        sum_alle = __reduce_add_sync(AVK_FULL, sum_alle);

        // ---- layout of the fixed part
        const int W = end - start;
        Npad = align_up(max(N, 1), 16);
        seq_cap = align_up(W + sum_l1 + 16, 16);
        wf_cap = align_up(2 * b0 + 3, 4);
        // room for equal-best results: all of them (<= max_branch_factor) when the arena is large,
        // a handful in the small shared-memory tiers (more than that escalates to the next tier)
        res_cap = (arena_bytes >= (256 << 10)) ? mbf : min(mbf, 8);
        long long off = region_off;
        vinfo = (VInfo *)(arena + off); off += (long long)sizeof(VInfo) * max(N, 1);
        u8 *alle_buf = arena + off;
        if (stage) off += align_up(sum_alle + 16, 16);
        bucket = (int *)(arena + off); off += align_up(4 * (N + 1), 16);
        res_alle = arena + off; off += (long long)res_cap * Npad;
        res_num = (int *)(arena + off); off += align_up(res_cap * 6 * 4, 16);
        hap_alle = arena + off; off += Npad;
        cur_obs = arena + off; off += 2 * Npad;
        best_obs = arena + off; off += 2 * Npad;
        if (off + 1024 > arena_bytes) return SOLVE_WORKSPACE;

        // ---- merged order: stable, truth before query on equal positions
        #pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            const int other = side ^ 1;
            #pragma unroll 1
            for (int i = lane; i < nv[side]; i += 32) {
                const u64 gv = v0[side] + i;
                const u32 p = b.pos[gv];
                int lo = 0, hi = nv[other];
                #pragma unroll 1
                while (lo < hi) {
                    const int m = (lo + hi) >> 1;
                    const u32 pm = b.pos[v0[other] + m];
                    const bool before = side == 0 ? (pm < p) : (pm <= p);
                    if (before) lo = m + 1; else hi = m;
                }
                VInfo v;
                v.pos = p; v.l0 = b.l0[gv]; v.l1 = b.l1[gv]; v.aoff = b.aoff[gv]; v.alt_ed = b.alt_ed[gv]; v.raw = b.raw[gv];
                v.gv = (u32)gv; v.type = b.vtype[gv]; v.zyg = b.zyg[gv]; v.is_truth = side == 0; v.slot = 0;
                vinfo[i + lo] = v;
            }
        }
        __syncwarp();
        // ---- stage allele bytes (shared-memory tiers) and assign metric-row slots
        alle_base = b.pool;
        if (stage) {
            int acc = 0;
            #pragma unroll 1
            for (int oi = 0; oi < N; ++oi) {
                const int n = (int)(vinfo[oi].l0 + vinfo[oi].l1);
                warp_copy(alle_buf + acc, b.pool + vinfo[oi].aoff, n);
                __syncwarp();
                if (lane == 0) vinfo[oi].aoff = (u32)acc;
                acc += n;
            }
            alle_base = alle_buf;
        }
        n_slots = 0;
        if (want_metrics) {
            u32 seen = 0;
            #pragma unroll 1
            for (int oi = 0; oi < N; ++oi) seen |= 1u << vinfo[oi].type;
            n_slots = __popc(seen);
            if (lane == 0) {
                int k = 0;
                #pragma unroll 1
                for (int t = 0; t < AVK_N_VARIANT_TYPES; ++t) if (seen & (1u << t)) slot_type[k++] = (u8)t;
                #pragma unroll 1
                for (int oi = 0; oi < N; ++oi) vinfo[oi].slot = (u8)__popc(seen & ((1u << vinfo[oi].type) - 1));
            }
            mrows = (u64 *)(arena + ((off + 7) / 8 * 8)); off = (off + 7) / 8 * 8 + 8LL * AVK_N_METRICS * (1 + n_slots);
            slot_tot = (u64 *)(arena + off); off += 16LL * max(n_slots, 1);
            slot_cnt = (u32 *)(arena + off); off += 16LL * max(n_slots, 1);
            if (off + 512 > arena_bytes) return SOLVE_WORKSPACE;
            #pragma unroll 1
            for (int i = lane; i < AVK_N_METRICS * (1 + n_slots); i += 32) mrows[i] = 0;
            #pragma unroll 1
            for (int i = lane; i < 2 * n_slots; i += 32) { slot_cnt[i] = 0; slot_tot[i] = 0; }
        }
        off = (off + 15) / 16 * 16;
        dyn = arena + off;
        dyn_bytes = arena_bytes - off;
        __syncwarp();
        return SOLVE_OK;
    }

    // Partition the dynamic part into queue arrays + node slots of `stride_` bytes.
    __device__ __noinline__ bool partition(int stride_, int min_slots) {
        stride = stride_;
        int ms = (int)min(dyn_bytes / (stride + 16), 60000LL);
        if (ms < min_slots) return false;
        long long off = 0;
        qkeys = (u64 *)(dyn + off); off += (long long)ms * 8;
        qslot = (u32 *)(dyn + off); off += (long long)ms * 4;
        freel = (u32 *)(dyn + off); off += (long long)ms * 4;
        off = (off + 15) / 16 * 16;
        if (off + (long long)ms * stride > dyn_bytes) ms = (int)((dyn_bytes - off) / stride);
        if (ms < min_slots) return false;
        max_slots = ms;
        nodes = dyn + off;
        #pragma unroll 1
        for (int i = lane_id(); i < ms; i += 32) freel[i] = (u32)(ms - 1 - i);
        nfree = ms;
        qn = 0;
        __syncwarp();
        return true;
    }
    __device__ __forceinline__ int alloc_slot() {   // warp-uniform; -1 when exhausted
        if (nfree == 0) return -1;
        nfree -= 1;
        return (int)freel[nfree];
    }
    __device__ __forceinline__ void free_slot(int s) {
        if (lane_id() == 0) freel[nfree] = (u32)s;
        nfree += 1;
        __syncwarp();
    }
    __device__ __forceinline__ void push(u64 key, int slot) {
        if (lane_id() == 0) { qkeys[qn] = key; qslot[qn] = (u32)slot; }
        qn += 1;
        __syncwarp();
    }
    // pop the minimum key: warp-parallel scan, then two REDUX min-reductions (high word, low word)
    // and a ballot to locate the owner -- keys are unique, so exactly one lane matches.
    __device__ __noinline__ int pop(u64 &key_out) {
        const int lane = lane_id();
        const int n = qn;
        u64 best = ~0ull;
        int bi = 0;
        #pragma unroll 1
        for (int i = lane; i < n; i += 32) { const u64 k = qkeys[i]; if (k < best) { best = k; bi = i; } }
        const u32 hi = (u32)(best >> 32), lo = (u32)best;
        const u32 mhi = __reduce_min_sync(AVK_FULL, hi);
        const u32 mlo = __reduce_min_sync(AVK_FULL, hi == mhi ? lo : 0xffffffffu);
        const int owner = __ffs(__ballot_sync(AVK_FULL, hi == mhi && lo == mlo)) - 1;
        bi = __shfl_sync(AVK_FULL, bi, owner);
        const int slot = (int)qslot[bi];
        __syncwarp();
        if (lane == 0) { qkeys[bi] = qkeys[n - 1]; qslot[bi] = qslot[n - 1]; }
        qn = n - 1;
        __syncwarp();
        key_out = ((u64)mhi << 32) | mlo;
        return slot;
    }

    // ------------------------------------------------------------------ tracker ops
    // copy_reference(): haplotype_dwfa.rs:218-227
    __device__ __forceinline__ bool copy_ref(u8 *seq, int &len, int &ref_pos, int to) {
        if (ref_pos < to) {
            const int n = to - ref_pos;
            if (len + n > seq_cap) return false;
            warp_copy(seq + len, ref + ref_pos, n);
            len += n;
            ref_pos = to;
        }
        return true;
    }
    // HaplotypeTracker::extend_variant(): haplotype_dwfa.rs:175-212.  false on capacity overflow.
    __device__ __noinline__ bool track_variant(u8 *seq, int &len, int &ref_pos, int &skip, const VInfo &v, bool alt, int sync,
                                                  bool &success) {
        const int vpos = (int)v.pos;
        if (!copy_ref(seq, len, ref_pos, vpos)) return false;
        success = true;
        if (alt) {
            if (ref_pos <= vpos) {
                const int l1 = (int)v.l1;
                if (len + l1 > seq_cap) return false;
                warp_copy(seq + len, alle_base + v.aoff + v.l0, l1);
                len += l1;
                ref_pos = vpos + (int)v.l0;
            } else {
                skip += (int)v.alt_ed;   // edit_distance(allele0, allele1) :199
                success = false;
            }
        }
        return copy_ref(seq, len, ref_pos, sync);
    }

    // ================================================================== optimize_sequences
    // node layout: int hdr[20] {id, depth, HapState h[2] (8 ints each), 2 pad}; u8 alle[Npad];
    //              u8 seq[4][seq_cap] (h0 truth, h0 query, h1 truth, h1 query); int wf[2][wf_cap]
    __device__ __forceinline__ int opt_stride() const { return 80 + Npad + 4 * seq_cap + 8 * wf_cap; }
    __device__ __forceinline__ int *n_hdr(int s) const { return (int *)(nodes + (long long)s * stride); }
    __device__ __forceinline__ u8 *n_alle(int s) const { return nodes + (long long)s * stride + 80; }
    __device__ __forceinline__ u8 *n_seq(int s, int k) const { return nodes + (long long)s * stride + 80 + Npad + (long long)k * seq_cap; }
    __device__ __forceinline__ int *n_wf(int s, int h) const { return (int *)(nodes + (long long)s * stride + 80 + Npad + 4LL * seq_cap) + (long long)h * wf_cap; }

    __device__ __noinline__ void opt_clone(int dst, int src) {
        const int lane = lane_id();
        int *hs = n_hdr(src), *hd = n_hdr(dst);
        const int depth = hs[1];
        const int tl0 = hs[2 + 2], ql0 = hs[2 + 3], ed0 = hs[2 + 6], tl1 = hs[10 + 2], ql1 = hs[10 + 3], ed1 = hs[10 + 6];
        if (lane < 20) hd[lane] = hs[lane];
        warp_copy(n_alle(dst), n_alle(src), depth);
        warp_copy(n_seq(dst, 0), n_seq(src, 0), tl0);
        warp_copy(n_seq(dst, 1), n_seq(src, 1), ql0);
        warp_copy(n_seq(dst, 2), n_seq(src, 2), tl1);
        warp_copy(n_seq(dst, 3), n_seq(src, 3), ql1);
        {
            const int *ws = n_wf(src, 0); int *wd = n_wf(dst, 0);
            #pragma unroll 1
            for (int i = lane; i < 2 * ed0 + 1; i += 32) wd[i] = ws[i];
            ws = n_wf(src, 1); wd = n_wf(dst, 1);
            #pragma unroll 1
            for (int i = lane; i < 2 * ed1 + 1; i += 32) wd[i] = ws[i];
        }
        __syncwarp();
    }

    // HaplotypeDWFA::extend_variant (haplotype_dwfa.rs:46-67) on hap h of node s.
    __device__ __noinline__ int opt_extend_hap(int s, int h, const VInfo &v, bool alt, int sync) {
        HapState *sp = (HapState *)(n_hdr(s) + 2 + 8 * h);
        HapState st = *sp;
        u8 *tseq = n_seq(s, 2 * h), *qseq = n_seq(s, 2 * h + 1);
        bool success;
        bool ok;
        if (v.is_truth) {
            ok = copy_ref(qseq, st.q_len, st.q_ref_pos, sync);
            ok = ok && track_variant(tseq, st.t_len, st.t_ref_pos, st.t_skip, v, alt, sync, success);
        } else {
            ok = copy_ref(tseq, st.t_len, st.t_ref_pos, sync);
            ok = ok && track_variant(qseq, st.q_len, st.q_ref_pos, st.q_skip, v, alt, sync, success);
        }
        if (!ok) return SOLVE_WORKSPACE;
        __syncwarp();
        int *wf = n_wf(s, h);
        const int rc = dwfa_update(wf, &st.ed, (wf_cap - 3) / 2, tseq, st.t_len, qseq, st.q_len, work);
        if (rc != DWFA_OK) return SOLVE_WORKSPACE;   // ED bound exceeded: never expected
        if (lane_id() == 0) *sp = st;
        __syncwarp();
        return SOLVE_OK;
    }
    __device__ __noinline__ int opt_extend(int s, int oi, bool a1_alt, bool a2_alt) {   // ComparisonNode::extend_variant :443-451
        const VInfo v = vinfo[oi];
        const int sync = sync_pos(oi);
        int rc = opt_extend_hap(s, 0, v, a1_alt, sync);
        if (rc) return rc;
        rc = opt_extend_hap(s, 1, v, a2_alt, sync);
        if (rc) return rc;
        if (lane_id() == 0) { n_alle(s)[oi] = (u8)((a1_alt ? 1 : 0) | (a2_alt ? 2 : 0)); n_hdr(s)[1] = oi + 1; }
        __syncwarp();
        return SOLVE_OK;
    }
    __device__ __forceinline__ u32 opt_cost(int s) const {
        const int *h = n_hdr(s);
        return (u32)(h[2 + 6] + h[2 + 4] + h[2 + 5] + h[10 + 6] + h[10 + 4] + h[10 + 5]);
    }
    __device__ __noinline__ int opt_finalize_hap(int s, int h) {   // finalize_dwfa: haplotype_dwfa.rs:84-95
        HapState *sp = (HapState *)(n_hdr(s) + 2 + 8 * h);
        HapState st = *sp;
        u8 *tseq = n_seq(s, 2 * h), *qseq = n_seq(s, 2 * h + 1);
        if (!copy_ref(tseq, st.t_len, st.t_ref_pos, end)) return SOLVE_WORKSPACE;
        if (!copy_ref(qseq, st.q_len, st.q_ref_pos, end)) return SOLVE_WORKSPACE;
        __syncwarp();
        int *wf = n_wf(s, h);
        const int cap = (wf_cap - 3) / 2;
        if (dwfa_update(wf, &st.ed, cap, tseq, st.t_len, qseq, st.q_len, work) != DWFA_OK) return SOLVE_WORKSPACE;
        if (dwfa_finalize(wf, &st.ed, cap, tseq, st.t_len, qseq, st.q_len, work) != DWFA_OK) return SOLVE_WORKSPACE;
        if (lane_id() == 0) *sp = st;
        __syncwarp();
        return SOLVE_OK;
    }

    // Runs the best-first search.  Results (all equal-best, in finalisation order) go to
    // res_alle/res_num.  stop_at_nonzero: merge only needs "is the minimum cost zero"
    // (merge_solver.rs:142-143); costs never decrease along a path, so the search may stop at
    // the first popped node whose cost is > 0.
    __device__ __noinline__ int optimize(bool stop_at_nonzero) {
        const int lane = lane_id();
        if (!partition(opt_stride(), min(N + 3, 48))) return SOLVE_WORKSPACE;   // cheap early escalation
        #pragma unroll 1
        for (int i = lane; i <= N; i += 32) bucket[i] = 0;
        n_res = 0;
        u32 best = 0xffffffffu;
        u32 next_id = 0;
        {   // root (query_optimizer.rs:184-192)
            const int s = alloc_slot();
            int *hdr = n_hdr(s);
            if (lane < 20) hdr[lane] = (lane == 2 || lane == 3 || lane == 10 || lane == 11) ? start : 0;
            if (lane == 0) { n_wf(s, 0)[0] = 0; n_wf(s, 1)[0] = 0; }
            __syncwarp();
            next_id = 1;
            push(0ull, s);
        }
        #pragma unroll 1
        while (qn > 0) {
            u64 key;
            const int s = pop(key);
            work.search_pops += 1;
            const u32 cost = (u32)(key >> 32);
            if (stop_at_nonzero && cost > 0) return SOLVE_OK;   // nothing cheaper is left; n_res tells if a zero-cost result exists
            if (cost > best) { free_slot(s); continue; }                       // :204 strict
            int *hdr = n_hdr(s);
            const int oi = hdr[1];
            const int bc = bucket[oi];
            if (bc >= mbf) { free_slot(s); continue; }                         // :222
            if (lane == 0) bucket[oi] = bc + 1;
            __syncwarp();
            if (oi == N) {                                                     // :227-247
                int rc = opt_finalize_hap(s, 0);
                if (rc) return rc;
                rc = opt_finalize_hap(s, 1);
                if (rc) return rc;
                const u32 c = opt_cost(s);
                if (c < best) { best = c; n_res = 0; }
                if (c == best && n_res >= res_cap) return SOLVE_WORKSPACE;
                if (c == best) {
                    warp_copy(res_alle + (long long)n_res * Npad, n_alle(s), N);
                    if (lane == 0) {
                        int *rn = res_num + n_res * 6;
                        rn[0] = hdr[2 + 6]; rn[1] = hdr[10 + 6]; rn[2] = hdr[2 + 4]; rn[3] = hdr[10 + 4]; rn[4] = hdr[2 + 5]; rn[5] = hdr[10 + 5];
                    }
                    __syncwarp();
                    n_res += 1;
                }
                free_slot(s);
                continue;
            }
            const int z = vinfo[oi].zyg;
            const bool is_truth = vinfo[oi].is_truth;
            const bool het = (z == AVK_ZYG_UNPHASED_HET || z == AVK_ZYG_PHASED_HET01 || z == AVK_ZYG_PHASED_HET10);
            if (het && (!is_truth || z == AVK_ZYG_UNPHASED_HET)) {             // :269-293
                const int s2 = alloc_slot();
                if (s2 < 0) return SOLVE_WORKSPACE;
                opt_clone(s2, s);
                // first child (REF, ALT) gets the lower id, second (ALT, REF) the next one
                if (lane == 0) { n_hdr(s2)[0] = (int)next_id; hdr[0] = (int)(next_id + 1); }
                __syncwarp();
                int rc = opt_extend(s2, oi, false, true);
                if (rc) return rc;
                push(((u64)opt_cost(s2) << 32) | next_id, s2);
                rc = opt_extend(s, oi, true, false);
                if (rc) return rc;
                push(((u64)opt_cost(s) << 32) | (next_id + 1), s);
                next_id += 2;
            } else if (het) {                                                  // :294-312 phased truth het, id kept
                const bool a1 = (z == AVK_ZYG_PHASED_HET10);
                const int rc = opt_extend(s, oi, a1, !a1);
                if (rc) return rc;
                push(((u64)opt_cost(s) << 32) | (u32)hdr[0], s);
            } else {
                if (z != AVK_ZYG_HOM_ALT) return AVK_ST_BAD_ZYGOSITY;          // assert_eq! :315
                const int rc = opt_extend(s, oi, true, true);
                if (rc) return rc;
                push(((u64)opt_cost(s) << 32) | (u32)hdr[0], s);
            }
        }
        if (n_res == 0) return AVK_ST_NO_RESULT;                               // :331
        return SOLVE_OK;
    }

    // ================================================================== optimize_gt_alleles
    // node layout: int hdr[8] {id, errors, depth, t_ref_pos, q_ref_pos, t_len, q_len, d}; u8 alle[Npad];
    //              u8 seq[2][seq_cap].  DWFA max ED is 0 (exact_gt_optimizer.rs:380), so the wavefront is the
    //              single matched length d and "alive" means one sequence is a prefix of the other.
    __device__ __forceinline__ int ex_stride() const { return 32 + Npad + 2 * seq_cap; }
    __device__ __forceinline__ u8 *x_alle(int s) const { return nodes + (long long)s * stride + 32; }
    __device__ __forceinline__ u8 *x_seq(int s, int k) const { return nodes + (long long)s * stride + 32 + Npad + (long long)k * seq_cap; }

    __device__ __noinline__ void ex_clone(int dst, int src) {
        const int lane = lane_id();
        int *hs = n_hdr(src), *hd = n_hdr(dst);
        const int depth = hs[2], tl = hs[5], ql = hs[6];
        if (lane < 8) hd[lane] = hs[lane];
        warp_copy(x_alle(dst), x_alle(src), depth);
        warp_copy(x_seq(dst, 0), x_seq(src, 0), tl);
        warp_copy(x_seq(dst, 1), x_seq(src, 1), ql);
        __syncwarp();
    }
    // ExactMatchNode::extend_variant (exact_gt_optimizer.rs:395-414): returns 1 keep / 0 drop / <0 error
    __device__ __noinline__ int ex_extend(int s, int oi, bool alt, bool is_error) {
        int *hdr = n_hdr(s);
        int t_ref_pos = hdr[3], q_ref_pos = hdr[4], t_len = hdr[5], q_len = hdr[6], d = hdr[7];
        const int errors = hdr[1];
        u8 *tseq = x_seq(s, 0), *qseq = x_seq(s, 1);
        const VInfo v = vinfo[oi];
        const int sync = sync_pos(oi);
        bool success, ok;
        int skip = 0;
        if (v.is_truth) {
            ok = copy_ref(qseq, q_len, q_ref_pos, sync);
            ok = ok && track_variant(tseq, t_len, t_ref_pos, skip, v, alt, sync, success);
        } else {
            ok = copy_ref(tseq, t_len, t_ref_pos, sync);
            ok = ok && track_variant(qseq, q_len, q_ref_pos, skip, v, alt, sync, success);
        }
        if (!ok) return SOLVE_WORKSPACE;
        __syncwarp();
        // DWFA update with max ED 0: extend the single diagonal, then require an end to be reached
        const int ext = warp_lcp(tseq + d, t_len - d, qseq + d, q_len - d);
        work.cells += 1; work.matched += (u32)ext;
        d += ext;
        const bool alive = (d >= t_len) || (d >= q_len);
        if (lane_id() == 0) {
            hdr[1] = errors + (is_error ? 1 : 0);
            hdr[2] = oi + 1;
            hdr[3] = t_ref_pos; hdr[4] = q_ref_pos; hdr[5] = t_len; hdr[6] = q_len; hdr[7] = d;
            x_alle(s)[oi] = alt ? AL_ALT : AL_REF;
        }
        __syncwarp();
        return (success && alive) ? 1 : 0;
    }
    __device__ __forceinline__ u64 ex_key(int s) const {   // (Reverse(errors), set - errors, Reverse(id)) :372,452-458
        const int *hdr = n_hdr(s);
        const u64 errors = (u64)hdr[1];
        const u64 good = (u64)(hdr[2] - hdr[1]);
        return (errors << 48) | ((0xffffull - good) << 32) | (u32)hdr[0];
    }

    // in: hap_alle[oi] (AL_REF / AL_ALT by order index).  out: obs[oi], *errors.
    __device__ __noinline__ int exact_gt(u8 *obs, int *errors_out) {
        const int lane = lane_id();
        if (N >= 0xffff) return SOLVE_WORKSPACE;
        if (!partition(ex_stride(), min(N + 3, 48))) return SOLVE_WORKSPACE;
        u32 next_id = 0;
        int best_err = 0x7fffffff;
        bool have_best = false;
        int min_sync = 0, af_index = 0, af_counts = 0;
        {
            const int s = alloc_slot();
            int *hdr = n_hdr(s);
            if (lane < 8) hdr[lane] = (lane == 3 || lane == 4) ? start : 0;
            __syncwarp();
            next_id = 1;
            push(ex_key(s), s);
        }
        #pragma unroll 1
        while (qn > 0) {
            u64 key;
            const int s = pop(key);
            work.exact_pops += 1;
            int *hdr = n_hdr(s);
            const int errors = hdr[1];
            if (errors >= best_err) { free_slot(s); continue; }                // :169 non-strict
            const int oi = hdr[2];
            if (oi == N) {                                                     // :180-192
                int t_ref_pos = hdr[3], q_ref_pos = hdr[4], t_len = hdr[5], q_len = hdr[6], d = hdr[7];
                u8 *tseq = x_seq(s, 0), *qseq = x_seq(s, 1);
                if (!copy_ref(tseq, t_len, t_ref_pos, end)) return SOLVE_WORKSPACE;
                if (!copy_ref(qseq, q_len, q_ref_pos, end)) return SOLVE_WORKSPACE;
                __syncwarp();
                const int ext = warp_lcp(tseq + d, t_len - d, qseq + d, q_len - d);
                work.cells += 1; work.matched += (u32)ext; work.alignments += 1;
                d += ext;
                const bool exact = (d >= t_len) && (d >= q_len);   // update ok + finalize ok <=> sequences equal
                if (exact && errors < best_err) {
                    best_err = errors;
                    have_best = true;
                    warp_copy(obs, x_alle(s), N);
                    __syncwarp();
                }
                free_slot(s);
                continue;
            }
            if (oi < min_sync) { free_slot(s); continue; }                     // :194-197
            if (hdr[5] == hdr[6] && hdr[3] == hdr[4]) {                        // is_synchronized (alive => ed == 0) :206-217
                min_sync = oi; af_counts = 0; af_index = oi;
            }
            const int al = hap_alle[oi];
            if (al == AL_REF) {                                                // :257-273 move, id kept
                const int keep = ex_extend(s, oi, false, false);
                if (keep < 0) return keep;
                if (keep) push(ex_key(s), s); else free_slot(s);
            } else {                                                           // :274-306 clone twice
                // (REF, error) first with id next_id; then (ALT, no error) unless auto-failed
                const bool do_alt = !(oi < af_index);
                int s_alt = -1;
                if (do_alt) {
                    s_alt = alloc_slot();
                    if (s_alt < 0) return SOLVE_WORKSPACE;
                    ex_clone(s_alt, s);
                }
                if (lane == 0) hdr[0] = (int)next_id;
                __syncwarp();
                next_id += 1;
                int keep = ex_extend(s, oi, false, true);
                if (keep < 0) return keep;
                if (keep) push(ex_key(s), s); else free_slot(s);
                if (do_alt) {
                    if (lane == 0) n_hdr(s_alt)[0] = (int)next_id;
                    __syncwarp();
                    next_id += 1;
                    keep = ex_extend(s_alt, oi, true, false);
                    if (keep < 0) return keep;
                    if (keep) push(ex_key(s_alt), s_alt); else free_slot(s_alt);
                }
            }
            af_counts += 1;                                                    // :310-339
            if (af_counts >= 500) {
                if (af_index >= N) return AVK_ST_NO_RESULT;   // the reference would index out of bounds here
                int w = 0;
                const int n = qn;
                #pragma unroll 1
                for (int i = 0; i < n; ++i) {
                    const int sl = (int)qslot[i];
                    const u8 a = x_alle(sl)[af_index];
                    const bool set = n_hdr(sl)[2] > af_index;
                    if (!set || a == AL_REF) {
                        if (lane == 0 && w != i) { qkeys[w] = qkeys[i]; qslot[w] = qslot[i]; }
                        w += 1;
                    } else {
                        free_slot(sl);
                    }
                    __syncwarp();
                }
                qn = w;
                af_index += 1;
                af_counts = 0;
            }
        }
        if (!have_best) return AVK_ST_NO_RESULT;                               // :345-348
        *errors_out = best_err;
        return SOLVE_OK;
    }

    // ================================================================== waffle solver
    // generate_allele_sequence(): waffle_solver.rs:726-778.  side 0 truth / 1 query, hap 0/1, alleles from
    // result r; type_filter < 0 keeps every variant.  Returns length (or -1), failed ED, #ALT spliced.
    __device__ __noinline__ int build_hap_seq(u8 *dst, int side, int hap, int r, int type_filter, int &failed, int &n_alt) {
        const u8 *ra = res_alle + (long long)r * Npad;
        int cur = start, len = 0;
        failed = 0; n_alt = 0;
        #pragma unroll 1
        for (int oi = 0; oi < N; ++oi) {
            const VInfo v = vinfo[oi];
            if ((v.is_truth ? 0 : 1) != side) continue;
            if (!((ra[oi] >> hap) & 1)) continue;                 // REF allele: skipped entirely (:738-741)
            if (type_filter >= 0 && v.type != type_filter) continue;
            const int vpos = (int)v.pos;
            if (vpos < cur) { failed += (int)v.alt_ed; continue; }   // :745-753
            if (!copy_ref(dst, len, cur, vpos)) return -1;
            const int l1 = (int)v.l1;
            if (len + l1 > seq_cap) return -1;
            warp_copy(dst + len, alle_base + v.aoff + v.l0, l1);
            len += l1;
            cur = vpos + (int)v.l0;
            n_alt += 1;
        }
        if (cur > end) return -2;
        if (!copy_ref(dst, len, cur, end)) return -1;
        __syncwarp();
        return len;
    }

    // GroupMetrics::add_truth_zygosity (grouped_metrics.rs:183-227); col 0 = truth columns, 2 = query columns
    // (the query pass lands in the query columns through add_swap_benchmark, :268-277).  lane 0 only.
    static __device__ __forceinline__ void gm_add(u64 *g, int col, u64 w, int exp, int obs) {
        if (exp == obs) {
            g[AVK_M_HAP + col] += exp; g[AVK_M_WEIGHTED_HAP + col] += exp * w; g[AVK_M_GT + col] += 1;
        } else {
            g[AVK_M_HAP + col] += obs; g[AVK_M_HAP + col + 1] += (exp - obs);
            g[AVK_M_WEIGHTED_HAP + col] += obs * w; g[AVK_M_WEIGHTED_HAP + col + 1] += (u64)(exp - obs) * w;
            g[AVK_M_GT + col + 1] += 1;
            if (obs > 0) g[(col == 0) ? AVK_M_GT_TRUTH_FN_GT : AVK_M_GT_QUERY_FP_GT] += 1;
        }
    }

    // global ED with overflow trap (the wavefront buffer is sized from the proven bound b0)
    __device__ __noinline__ int ed_checked(bool &ovf, const u8 *A, int la, const u8 *B, int lb, int *wf, int cap) {
        const int e = wfa_ed_warp(A, la, B, lb, wf, cap, work);
        if (e < 0) { ovf = true; return 0; }
        return e;
    }

    __device__ int solve_compare(u64 r, const avk_compare_cfg &cfg, const DevCompareOut &out);
    __device__ int solve_merge(u64 r, const avk_merge_cfg &cfg, const DevMergeOut &out);
};

// solve_compare_region(): returns AVK_ST_* (>= 0) or SOLVE_WORKSPACE
__device__ int RegionSolver::solve_compare(u64 r, const avk_compare_cfg &cfg, const DevCompareOut &out) {
    const DevBatch &b = *bp;
    const int lane = lane_id();
    const u32 c = b.contig[r];
    start = (int)b.start[r];
    end = (int)b.end[r];
    if (c >= b.n_contigs || b.start[r] > b.end[r] || (u64)b.end[r] > b.contig_len[c] || b.end[r] > 0x7fff0000u) return AVK_ST_BAD_INPUT;
    mbf = (int)cfg.max_branch_factor;
    if (mbf <= 0) return AVK_ST_BAD_INPUT;
    if (!begin_region(b.contig_ptr[c])) return SOLVE_WORKSPACE;
    bool bad;
    int rc = setup_pair(r, 0, 1, true, bad);
    if (rc) return rc;
    if (bad) return AVK_ST_BAD_INPUT;
    const int W = end - start;

    rc = optimize(false);
    if (rc) return rc;

    int best_r = 0;
    const bool shortcut = cfg.enable_exact_shortcut &&
        (res_num[0] + res_num[1] + res_num[2] + res_num[3] + res_num[4] + res_num[5] == 0);   // :171
    if (!shortcut) {
        // ---- exact-GT scoring of every equal-best solution; first minimum wins (:169-265)
        int best_total = 0x7fffffff;
        #pragma unroll 1
        for (int ri = 0; ri < n_res; ++ri) {
            int total = 0;
            #pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const u8 *ra = res_alle + (long long)ri * Npad;
                #pragma unroll 1
                for (int i = lane; i < N; i += 32) hap_alle[i] = ((ra[i] >> h) & 1) ? AL_ALT : AL_REF;
                __syncwarp();
                int errs = 0;
                const int *rnum = res_num + ri * 6;
                if (rnum[h] + rnum[2 + h] + rnum[4 + h] == 0) {
                    // ED 0 and nothing skipped on this haplotype: the zero-flip path of optimize_gt_alleles
                    // replays exactly these tracker steps, stays alive, is popped first (fewest errors, most
                    // set alleles) and finalises with 0 errors, after which every other node is pruned
                    // (errors >= best, exact_gt_optimizer.rs:169).  Its result is the input alleles.
                    warp_copy(cur_obs + h * Npad, hap_alle, N);
                    __syncwarp();
                } else {
                    rc = exact_gt(cur_obs + h * Npad, &errs);
                    if (rc) return rc;
                }
                total += errs;
            }
            if (total < best_total) {
                best_total = total;
                best_r = ri;
                warp_copy(best_obs, cur_obs, 2 * Npad);
                __syncwarp();
            }
        }
    }
    const u8 *ra = res_alle + (long long)best_r * Npad;
    u64 *gm = mrows;   // joint row; type rows follow at gm + 22 * (1 + slot)

    // ---- per-variant expected/observed (:226-258); lanes across variants
    #pragma unroll 1
    for (int oi = lane; oi < N; oi += 32) {
        const VInfo v = vinfo[oi];
        int exp_, obs_;
        if (shortcut) {   // generate_exact_match(): counts come from the RAW zygosities (:544-557)
            exp_ = (v.zyg == AVK_ZYG_HOM_ALT) ? 2 : 1;
            obs_ = exp_;
        } else {
            exp_ = (ra[oi] & 1) + ((ra[oi] >> 1) & 1);
            obs_ = (best_obs[oi] == AL_ALT ? 1 : 0) + (best_obs[Npad + oi] == AL_ALT ? 1 : 0);
        }
        const int cls = (exp_ == obs_) ? AVK_CLASS_TP : (v.is_truth ? AVK_CLASS_FN : AVK_CLASS_FP);
        out.vexp[v.gv] = (u8)(v.is_truth ? exp_ : obs_);     // query entries are toggled (compare_benchmark.rs:109-123)
        out.vobs[v.gv] = (u8)(v.is_truth ? obs_ : exp_);
        out.vcls[v.gv] = (u8)cls;
        hap_alle[oi] = (u8)(exp_ | (obs_ << 4));             // reuse as scratch for the metric pass below
    }
    __syncwarp();
    // ---- GT / HAP / WEIGHTED_HAP (+ add_swap_benchmark :269), accumulated in the arena by lane 0
    if (lane == 0) {
        #pragma unroll 1
        for (int oi = 0; oi < N; ++oi) {
            const VInfo v = vinfo[oi];
            const int exp_ = hap_alle[oi] & 15, obs_ = hap_alle[oi] >> 4;
            const int col = v.is_truth ? 0 : 2;
            gm_add(gm, col, v.alt_ed, exp_, obs_);
            gm_add(gm + AVK_N_METRICS * (1 + v.slot), col, v.alt_ed, exp_, obs_);
            slot_cnt[2 * v.slot + (v.is_truth ? 0 : 1)] += 1;
        }
    }
    __syncwarp();

    // ---- basepair metrics: three sequence buffers + one wavefront in the dynamic part
    const long long need = 3LL * seq_cap + 4LL * wf_cap + 16;
    if (need > dyn_bytes) return SOLVE_WORKSPACE;
    u8 *bufT = dyn, *bufQ = dyn + seq_cap, *bufF = dyn + 2LL * seq_cap;
    int *wf = (int *)(dyn + 3LL * seq_cap);
    const int ed_cap = (wf_cap - 3) / 2;
    const u8 *R = ref + start;
    const bool want_seq = out.seq_off && cfg.enable_sequences;
    bool ed_overflow = false;
    u32 mask = 0;
    #pragma unroll 1
    for (int k = 0; k < n_slots; ++k) mask |= 1u << slot_type[k];

    int status = AVK_ST_OK;
    if (shortcut) {
        // generate_exact_match(): :559-598
        u64 js = 0;
        #pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            int failed, n_alt;
            const int lt = build_hap_seq(bufT, 0, h, 0, -1, failed, n_alt);
            if (lt < 0) return lt == -1 ? SOLVE_WORKSPACE : AVK_ST_BAD_INPUT;
            js += (u64)ed_checked(ed_overflow, R, W, bufT, lt, wf, ed_cap);
            if (want_seq) {
                const int lq = build_hap_seq(bufQ, 1, h, 0, -1, failed, n_alt);
                if (lq < 0) return lq == -1 ? SOLVE_WORKSPACE : AVK_ST_BAD_INPUT;
                warp_copy(out.seq_pool + out.seq_off[r * 5 + 1 + h], bufT, lt);
                warp_copy(out.seq_pool + out.seq_off[r * 5 + 3 + h], bufQ, lq);
                if (lane == 0) { out.seq_len[r * 5 + 1 + h] = (u32)lt; out.seq_len[r * 5 + 3 + h] = (u32)lq; }
            }
        }
        if (lane == 0) {
            gm[AVK_M_BASEPAIR + 0] += 2 * js; gm[AVK_M_BASEPAIR + 2] += 2 * js;
            #pragma unroll 1
            for (int oi = 0; oi < N; ++oi) {
                const VInfo v = vinfo[oi];
                const u64 dd = 2ull * v.alt_ed * (u64)((v.zyg == AVK_ZYG_HOM_ALT) ? 2 : 1);
                gm[AVK_N_METRICS * (1 + v.slot) + AVK_M_BASEPAIR + (v.is_truth ? 0 : 2)] += dd;
            }
        }
    } else {
        // add_basepair_stats(): :335-449
        #pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            int failT, failQ, altT, altQ;
            const int lt = build_hap_seq(bufT, 0, h, best_r, -1, failT, altT);
            if (lt < 0) return lt == -1 ? SOLVE_WORKSPACE : AVK_ST_BAD_INPUT;
            const int lq = build_hap_seq(bufQ, 1, h, best_r, -1, failQ, altQ);
            if (lq < 0) return lq == -1 ? SOLVE_WORKSPACE : AVK_ST_BAD_INPUT;
            if (want_seq) {
                warp_copy(out.seq_pool + out.seq_off[r * 5 + 1 + h], bufT, lt);
                warp_copy(out.seq_pool + out.seq_off[r * 5 + 3 + h], bufQ, lq);
                if (lane == 0) { out.seq_len[r * 5 + 1 + h] = (u32)lt; out.seq_len[r * 5 + 3 + h] = (u32)lq; }
            }
            // X = ED(ref, truth), Y = ED(ref, query), Z = ED(truth, query); a haplotype without a
            // spliced ALT IS the reference window, so its distances are known without aligning.
            u64 X, Y, Z;
            const int *rnb = res_num + best_r * 6;
            X = altT ? (u64)ed_checked(ed_overflow, R, W, bufT, lt, wf, ed_cap) : 0;
            if (rnb[h] == 0) {          // optimizer ED(truth, query) == 0 on this haplotype: identical sequences
                Y = X; Z = 0;
            } else {
                Y = altQ ? (u64)ed_checked(ed_overflow, R, W, bufQ, lq, wf, ed_cap) : 0;
                if (!altT) Z = Y; else if (!altQ) Z = X; else Z = (u64)rnb[h];   // Z is the optimizer's finalised ED
            }
            const u64 tp = X + Y - Z;                    // (2X + 2Y - 2Z) / 2  :644
            if (lane == 0) {
                gm[AVK_M_BASEPAIR + 0] += tp; gm[AVK_M_BASEPAIR + 1] += 2 * X - tp + 2 * (u64)failT;
                gm[AVK_M_BASEPAIR + 2] += tp; gm[AVK_M_BASEPAIR + 3] += 2 * Y - tp + 2 * (u64)failQ;
            }
            // per supported type that occurs in the cluster (absent types only ever receive zeros, :395-444)
            #pragma unroll 1
            for (int k = 0; k < n_slots; ++k) {
                const int ft = slot_type[k];
                if (!type_supported(ft)) continue;
                const int ntf = (int)slot_cnt[2 * k], nqf = (int)slot_cnt[2 * k + 1];
                u64 q_tp = 0, q_fp = 0, t_tp = 0, t_fn = 0;
                if (nqf > 0) {                                           // :395-410
                    if (nqf == nv[1]) { q_tp = tp; q_fp = 2 * Y - tp + 2 * (u64)failQ; }   // filtered == full query
                    else {
                        int failF, altF;
                        const int lf = build_hap_seq(bufF, 1, h, best_r, ft, failF, altF);
                        if (lf < 0) return lf == -1 ? SOLVE_WORKSPACE : AVK_ST_BAD_INPUT;
                        u64 Yf, Zf;
                        if (!altF) { Yf = 0; Zf = X; }
                        else {
                            Yf = (u64)ed_checked(ed_overflow, R, W, bufF, lf, wf, ed_cap);
                            Zf = altT ? (u64)ed_checked(ed_overflow, bufT, lt, bufF, lf, wf, ed_cap) : Yf;
                        }
                        const u64 tpf = X + Yf - Zf;
                        q_tp = tpf; q_fp = 2 * Yf - tpf + 2 * (u64)failF;
                    }
                }
                if (ntf > 0) {                                           // :422-437
                    if (ntf == nv[0]) { t_tp = tp; t_fn = 2 * X - tp + 2 * (u64)failT; }
                    else {
                        int failF, altF;
                        const int lf = build_hap_seq(bufF, 0, h, best_r, ft, failF, altF);
                        if (lf < 0) return lf == -1 ? SOLVE_WORKSPACE : AVK_ST_BAD_INPUT;
                        u64 Xf, Zf;
                        if (!altF) { Xf = 0; Zf = Y; }
                        else {
                            Xf = (u64)ed_checked(ed_overflow, R, W, bufF, lf, wf, ed_cap);
                            Zf = altQ ? (u64)ed_checked(ed_overflow, bufF, lf, bufQ, lq, wf, ed_cap) : Xf;
                        }
                        const u64 tpf = Xf + Y - Zf;
                        t_tp = tpf; t_fn = 2 * Xf - tpf + 2 * (u64)failF;
                    }
                }
                if (lane == 0) {
                    u64 *g = gm + AVK_N_METRICS * (1 + k) + AVK_M_BASEPAIR;
                    g[0] += t_tp; g[1] += t_fn; g[2] += q_tp; g[3] += q_fp;
                }
            }
        }
        // every supported type gets a (possibly all-zero) entry (:444)
        mask |= (1u << AVK_VT_SNV) | (1u << AVK_VT_INSERTION) | (1u << AVK_VT_DELETION) | (1u << AVK_VT_INDEL) |
                (1u << AVK_VT_TR_CONTRACTION) | (1u << AVK_VT_TR_EXPANSION) | (1u << AVK_VT_SV_DELETION) | (1u << AVK_VT_SV_INSERTION);
        __syncwarp();
        // add_record_basepair_stats(): :455-522 (wrapping u64 like a release build).  Types without
        // variants have zero totals and zero basepair counts, so their record rows stay zero.
        if (lane == 0) {
            u64 truth_total = 0, query_total = 0;
            u64 *tq = slot_tot;
            #pragma unroll 1
            for (int oi = 0; oi < N; ++oi) {
                const VInfo v = vinfo[oi];
                const u64 cnt = (u64)((v.zyg == AVK_ZYG_HOM_ALT) ? 2 : 1) * v.raw;
                tq[2 * v.slot + (v.is_truth ? 0 : 1)] += cnt;
                if (v.is_truth) truth_total += cnt; else query_total += cnt;
            }
            const u64 tfn = gm[AVK_M_BASEPAIR + 1], qfp = gm[AVK_M_BASEPAIR + 3];
            const u64 ttp = 2 * truth_total - tfn, qtp = 2 * query_total - qfp;
            if (!(ttp >= gm[AVK_M_BASEPAIR + 0]) || !(qtp >= gm[AVK_M_BASEPAIR + 2])) status = AVK_ST_TP_UNDERFLOW;
            else {
                gm[AVK_M_RECORD_BP + 0] += ttp; gm[AVK_M_RECORD_BP + 1] += tfn; gm[AVK_M_RECORD_BP + 2] += qtp; gm[AVK_M_RECORD_BP + 3] += qfp;
                #pragma unroll 1
                for (int k = 0; k < n_slots; ++k) {
                    u64 *g = gm + AVK_N_METRICS * (1 + k);
                    const u64 fn_ = g[AVK_M_BASEPAIR + 1], fp_ = g[AVK_M_BASEPAIR + 3];
                    g[AVK_M_RECORD_BP + 0] += 2 * tq[2 * k] - fn_; g[AVK_M_RECORD_BP + 1] += fn_;
                    g[AVK_M_RECORD_BP + 2] += 2 * tq[2 * k + 1] - fp_; g[AVK_M_RECORD_BP + 3] += fp_;
                }
            }
        }
        status = __shfl_sync(AVK_FULL, status, 0);
    }
    if (ed_overflow) return SOLVE_WORKSPACE;
    if (status != AVK_ST_OK) return status;
    __syncwarp();
    // ---- write the full GroupTypeMetrics row [13][22] (coalesced), zeros for absent types
    {
        u64 *dst = out.region_metrics + r * (u64)(AVK_N_GROUPS * AVK_N_METRICS);
        // type -> slot lookup in a register: 4 bits per type, 15 = absent
        u64 lut = ~0ull;
        #pragma unroll 1
        for (int k = 0; k < n_slots; ++k) lut = (lut & ~(15ull << (4 * slot_type[k]))) | ((u64)k << (4 * slot_type[k]));
        #pragma unroll 1
        for (int i = lane; i < AVK_N_GROUPS * AVK_N_METRICS; i += 32) {
            const int g = i / AVK_N_METRICS, m = i - g * AVK_N_METRICS;
            u64 v = 0;
            if (g == 0) v = gm[m];
            else {
                const int k = (int)((lut >> (4 * (g - 1))) & 15);
                if (k != 15) v = gm[AVK_N_METRICS * (1 + k) + m];
            }
            dst[i] = v;
        }
    }
    if (want_seq) {
        warp_copy(out.seq_pool + out.seq_off[r * 5], R, W);
        if (lane == 0) out.seq_len[r * 5] = (u32)W;
    }
    if (lane == 0) {
        const int *rn = res_num + best_r * 6;
        out.ed1[r] = shortcut ? 0u : (u32)rn[0];
        out.ed2[r] = shortcut ? 0u : (u32)rn[1];
        out.type_mask[r] = (uint16_t)mask;
    }
    return AVK_ST_OK;
}

// solve_merge_region(): merge_solver.rs:110-200
__device__ int RegionSolver::solve_merge(u64 r, const avk_merge_cfg &cfg, const DevMergeOut &out) {
    const DevBatch &b = *bp;
    const int lane = lane_id();
    const u32 K = b.n_inputs;
    const u32 c = b.contig[r];
    start = (int)b.start[r];
    end = (int)b.end[r];
    if (c >= b.n_contigs || b.start[r] > b.end[r] || (u64)b.end[r] > b.contig_len[c] || b.end[r] > 0x7fff0000u) return AVK_ST_BAD_INPUT;
    mbf = (int)cfg.max_branch_factor;
    if (mbf <= 0 || K > 32) return AVK_ST_BAD_INPUT;
    {
        bool invalid = false;
        int d0 = 0, d1 = 0, d2 = 0;
        #pragma unroll 1
        for (u32 k = 0; k < K; ++k) {
            const u64 v0 = b.var_off[r * K + k];
            invalid = validate_list(v0, (int)(b.var_off[r * K + k + 1] - v0), d0, d1, d2) || invalid;
        }
        if (__any_sync(AVK_FULL, invalid)) return AVK_ST_BAD_INPUT;
    }
    if (!begin_region(b.contig_ptr[c])) return SOLVE_WORKSPACE;
    // variant_delta_length(): :211-223, lane k holds input k
    long long delta = 0;
    bool unknown = false;
    if ((u32)lane < K) {
        #pragma unroll 1
        for (u64 gv = b.var_off[r * K + lane]; gv < b.var_off[r * K + lane + 1]; ++gv) {
            const int z = b.zyg[gv];
            unknown = unknown || z == AVK_ZYG_UNKNOWN;
            const long long cnt = (z == AVK_ZYG_HOM_ALT) ? 2 : ((z == AVK_ZYG_UNPHASED_HET || z == AVK_ZYG_PHASED_HET01 || z == AVK_ZYG_PHASED_HET10) ? 1 : 0);
            delta += ((long long)b.l1[gv] - (long long)b.l0[gv]) * cnt;
        }
    }
    if (__any_sync(AVK_FULL, unknown)) return AVK_ST_BAD_ZYGOSITY;
    u32 match_row = ((u32)lane < K) ? (1u << lane) : 0;   // lane i holds match_sets[i] as a bit mask
    bool all_identical = true, no_conflict = true;
    #pragma unroll 1
    for (u32 i = 0; i < K; ++i) {
        #pragma unroll 1
        for (u32 j = i + 1; j < K; ++j) {
            const long long di = __shfl_sync(AVK_FULL, delta, i), dj = __shfl_sync(AVK_FULL, delta, j);
            bool exact = false;
            const bool empty_i = b.var_off[r * K + i + 1] == b.var_off[r * K + i];
            const bool empty_j = b.var_off[r * K + j + 1] == b.var_off[r * K + j];
            if (di == dj) {                                              // :135-143
                bool bad;
                int rc = setup_pair(r, i, j, false, bad);
                if (rc) return rc;
                if (bad) return AVK_ST_BAD_INPUT;
                rc = optimize(true);
                if (rc) return rc;
                exact = n_res > 0 && (res_num[0] + res_num[1] + res_num[2] + res_num[3] + res_num[4] + res_num[5] == 0);
            }
            all_identical = all_identical && exact;
            no_conflict = no_conflict && (empty_i || empty_j || exact);   // :155-157
            if (exact) {
                if ((u32)lane == i) match_row |= 1u << j;
                if ((u32)lane == j) match_row |= 1u << i;
            }
        }
    }
    const u32 maj = K / 2 + 1;                                            // :167
    const unsigned has = __ballot_sync(AVK_FULL, (u32)lane < K && (u32)__popc(match_row) >= maj);
    u32 first_maj = 0;
    if (has) first_maj = __shfl_sync(AVK_FULL, match_row, __ffs(has) - 1);
    int cls;
    u32 idx_mask = 0;
    int sel = -1;
    if (all_identical) cls = AVK_MERGE_BASEPAIR_IDENTICAL;               // :174-197
    else if (cfg.no_conflict_enabled && no_conflict) {
        cls = AVK_MERGE_NO_CONFLICT;
        const bool nonempty = (u32)lane < K && b.var_off[r * K + lane + 1] != b.var_off[r * K + lane];
        idx_mask = __ballot_sync(AVK_FULL, nonempty);
    } else if (cfg.majority_voting_enabled && first_maj != 0) { cls = AVK_MERGE_MAJORITY_AGREE; idx_mask = first_maj; }
    else if (cfg.conflict_selection >= 0) { cls = AVK_MERGE_CONFLICT_SELECTION; sel = cfg.conflict_selection; }
    else cls = AVK_MERGE_DIFFERENT;
    if (lane == 0) {
        out.cls[r] = (u8)cls;
        int n = 0;
        if (sel >= 0) { out.idx[r * K + 0] = (u8)sel; n = 1; }
        else for (u32 k = 0; k < K; ++k) if (idx_mask & (1u << k)) out.idx[r * K + (n++)] = (u8)k;
        #pragma unroll 1
        for (u32 k = n; k < K; ++k) out.idx[r * K + k] = 0xFF;
        out.n_idx[r] = (u8)n;
    }
    return AVK_ST_OK;
}

}  // namespace avk
