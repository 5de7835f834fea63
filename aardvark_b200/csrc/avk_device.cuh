// avk_device.cuh -- warp-cooperative device primitives of the haplotype-comparison path.
//
// One warp owns one cluster (or one alignment).  All 32 lanes execute these
// functions together with warp-uniform scalar arguments; lanes split either
//   * the BASES of one diagonal (4 bytes per lane, 128 bytes per warp step:
//     unaligned 32-bit loads built with funnel shifts, XOR, __ffs, __ballot_sync), or
//   * the DIAGONALS of one wavefront (one diagonal per lane),
// whichever the wavefront width calls for.  Nothing here uses tensor cores:
// this is integer DP, bounded by the INT32 ALU pipe and latency.
//
// Sequences are raw bytes (the reference compares raw bytes, dynamic_wfa.rs:118, so
// N / IUPAC / soft-masked bases stay distinct symbols).  Every sequence buffer the
// primitives read has >= 8 readable bytes of slack after its logical end.
//
// Semantics follow the reference exactly (paths relative to the reference repo):
//   DWFALite::{extend,increase_edit_distance,update,finalize}  src/dwfa/dynamic_wfa.rs:68-245
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define AVK_FULL 0xffffffffu

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;

namespace avk {

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- work counters (DESIGN.md "algorithmic work"), per warp, flushed once at kernel end -----------
struct WorkAcc {
    u64 cells, matched;
    u32 alignments, search_pops, exact_pops;
    __device__ __forceinline__ void clear() { alignments = cells = matched = search_pops = exact_pops = 0; }
};

// ---- unaligned 32-bit load: two aligned loads + funnel shift ---------------------------------
__device__ __forceinline__ u32 ld4u(const u8 *p) {
    const uintptr_t a = (uintptr_t)p;
    const u32 *w = (const u32 *)(a & ~(uintptr_t)3);
    const u32 sh = ((u32)a & 3u) * 8u;
    return __funnelshift_r(w[0], w[1], sh);   // sh == 0 returns w[0]; w[1] is inside the slack
}

// ---- byte copy, any alignment, generic pointers (shared or global) ---------------------------
__device__ __noinline__ void warp_copy(u8 *dst, const u8 *src, int n) {
    const int lane = lane_id();
    if (n < 16) {
        if (lane < n) dst[lane] = src[lane];
        return;
    }
    const int head = (int)((4 - ((uintptr_t)dst & 3)) & 3);
    if (lane < head) dst[lane] = src[lane];
    u32 *d4 = (u32 *)(dst + head);
    const u8 *s = src + head;
    const int nw = (n - head) >> 2;
    #pragma unroll 1
    for (int i = lane; i < nw; i += 32) d4[i] = ld4u(s + 4 * i);
    const int done = head + (nw << 2);
    if (done + lane < n) dst[done + lane] = src[done + lane];
}

// ---- longest common prefix, lanes across bases (128 bytes per step) --------------------------
// number of equal leading bytes of a[0..na) and b[0..nb); warp-uniform result.
__device__ __noinline__ int warp_lcp(const u8 *a, int na, const u8 *b, int nb) {
    const int lane = lane_id();
    const int maxn = min(na, nb);
    int total = 0;
    #pragma unroll 1
    while (total < maxn) {
        const int k = total + 4 * lane;
        int good = 0;   // equal bytes in this lane's word, capped by the bytes that exist
        const int valid = min(4, maxn - k);
        if (valid > 0) {
            const u32 x = ld4u(a + k) ^ ld4u(b + k);
            good = x ? ((__ffs(x) - 1) >> 3) : 4;
            good = min(good, valid);
        }
        const unsigned m = __ballot_sync(AVK_FULL, good == 4);
        if (m == AVK_FULL) { total += 128; continue; }
        const int f = __ffs(~m) - 1;
        total += 4 * f + __shfl_sync(AVK_FULL, good, f);
        break;
    }
    return min(total, maxn);
}

// ---- DWFA -----------------------------------------------------------------------------------
// State: ed, wavefront wf[0 .. 2*ed] (int32, bases consumed in `other`), kept by the caller.
// baseline offset of entry i is wf[i] + ed - i (dynamic_wfa.rs:114).
struct Reach {
    int max_base;   // max_i (wf[i] + ed - i)     (maximum_baseline_distance :201-208)
    int max_other;  // max_i wf[i]                (maximum_other_distance    :212-215)
    bool full;      // any i: base >= la && other >= lb (reached_full_diagonal :237-245)
};

// extend(): dynamic_wfa.rs:94-130.  Narrow wavefronts: one warp-wide LCP per diagonal.
// Wide wavefronts: one diagonal per lane, one word compare each; lanes whose first word matched
// completely are finished with a warp-wide LCP so that one long run does not serialise the warp.
__device__ __noinline__ Reach dwfa_extend(int *wf, int ed, const u8 *A, int la, const u8 *B, int lb, WorkAcc &w) {
    const int lane = lane_id();
    const int n = 2 * ed + 1;
    int mb = -1, mo = -1;
    bool full = false;
    int matched = 0;
    if (n <= 3) {
        #pragma unroll 1
        for (int i = 0; i < n; ++i) {
            int d = wf[i];
            int boff = d + ed - i;
            int ext = 0;
            if (boff < la && d < lb) ext = warp_lcp(A + boff, la - boff, B + d, lb - d);
            d += ext; boff += ext; matched += ext;
            if (ext && lane == 0) wf[i] = d;
            mb = max(mb, boff); mo = max(mo, d);
            full = full || (boff >= la && d >= lb);
        }
    } else {
        #pragma unroll 1
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const bool act = i < n;
            int d = act ? wf[i] : 0;
            int boff = d + ed - i;
            const int d0 = d;
            bool more = false;
            if (act && boff < la && d < lb) {
                const int valid = min(4, min(la - boff, lb - d));
                const u32 x = ld4u(A + boff) ^ ld4u(B + d);
                int good = x ? ((__ffs(x) - 1) >> 3) : 4;
                good = min(good, valid);
                d += good; boff += good;
                more = (good == 4) && boff < la && d < lb;
            }
            unsigned m = __ballot_sync(AVK_FULL, more);
            #pragma unroll 1
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int dd = __shfl_sync(AVK_FULL, d, src);
                const int bb = __shfl_sync(AVK_FULL, boff, src);
                const int ext = warp_lcp(A + bb, la - bb, B + dd, lb - dd);
                if (lane == src) { d += ext; boff += ext; }
            }
            if (act) {
                if (d != d0) wf[i] = d;
                matched += d - d0;
                mb = max(mb, boff); mo = max(mo, d);
                full = full || (boff >= la && d >= lb);
            }
        }
        mb = __reduce_max_sync(AVK_FULL, mb);
        mo = __reduce_max_sync(AVK_FULL, mo);
        full = __any_sync(AVK_FULL, full);
        matched = __reduce_add_sync(AVK_FULL, matched);
    }
    __syncwarp();
    w.cells += (u32)n;
    w.matched += (u32)matched;
    Reach r;
    r.max_base = mb; r.max_other = mo; r.full = full;
    return r;
}

// increase_edit_distance() without the re-extend: dynamic_wfa.rs:152-168, in place.
// new[i] = max(old[i], old[i-1]+1, old[i-2]+1) over the entries that exist; chunks are
// processed from the top so that every read of old[] precedes the write that replaces it.
__device__ __noinline__ void dwfa_grow(int *wf, int old_ed) {
    const int lane = lane_id();
    const int n_old = 2 * old_ed + 1;
    const int n_new = n_old + 2;
    #pragma unroll 1
    for (int base = ((n_new - 1) >> 5) << 5; base >= 0; base -= 32) {
        const int i = base + lane;
        int v = 0;
        if (i < n_new) {
            if (i < n_old) v = wf[i];
            if (i >= 1 && i - 1 < n_old) v = max(v, wf[i - 1] + 1);
            if (i >= 2 && i - 2 < n_old) v = max(v, wf[i - 2] + 1);
        }
        __syncwarp();
        if (i < n_new) wf[i] = v;
    }
    __syncwarp();
}

enum { DWFA_OK = 0, DWFA_MAX_ED = 1 };

// update(): dynamic_wfa.rs:68-84.  *ed is left incremented when the cap is hit (:146-149).
__device__ __noinline__ int dwfa_update(int *wf, int *ed, int max_ed, const u8 *A, int la, const u8 *B, int lb, WorkAcc &w) {
    Reach r = dwfa_extend(wf, *ed, A, la, B, lb, w);
    #pragma unroll 1
    while (!(r.max_base >= la) && !(r.max_other >= lb)) {
        *ed += 1;
        if (*ed > max_ed) return DWFA_MAX_ED;
        dwfa_grow(wf, *ed - 1);
        r = dwfa_extend(wf, *ed, A, la, B, lb, w);
    }
    return DWFA_OK;
}

// finalize(): dynamic_wfa.rs:183-198
__device__ __noinline__ int dwfa_finalize(int *wf, int *ed, int max_ed, const u8 *A, int la, const u8 *B, int lb, WorkAcc &w) {
    w.alignments += 1;
    Reach r = dwfa_extend(wf, *ed, A, la, B, lb, w);
    #pragma unroll 1
    while (!r.full) {
        *ed += 1;
        if (*ed > max_ed) return DWFA_MAX_ED;
        dwfa_grow(wf, *ed - 1);
        r = dwfa_extend(wf, *ed, A, la, B, lb, w);
    }
    return DWFA_OK;
}

// wfa_ed(): src/util/sequence_alignment.rs:9-13.  wf holds 2*max_ed+3 ints; returns -1 if the
// distance would exceed max_ed (callers size the buffer from a proven bound, so -1 is a bug trap).
__device__ __noinline__ int wfa_ed_warp(const u8 *A, int la, const u8 *B, int lb, int *wf, int max_ed, WorkAcc &w) {
    if (lane_id() == 0) wf[0] = 0;
    __syncwarp();
    int ed = 0;
    if (dwfa_finalize(wf, &ed, max_ed, A, la, B, lb, w) != DWFA_OK) return -1;
    return ed;
}

// ---- TMA bulk copy of a reference window into shared memory ------------------------------------
// cp.async.bulk (1-D bulk tensor-less TMA, SASS UBLKCP) + mbarrier completion.  src/dst 16-byte
// aligned, bytes a multiple of 16.  One mbarrier per warp; `phase` toggles per use.
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *mbar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __noinline__ void tma_window_load(u8 *smem_dst, const u8 *gsrc, u32 bytes, u64 *mbar, u32 &phase) {
    // order this warp's earlier generic-proxy accesses to the window before the async-proxy write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane_id() == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
    }
    u32 done = 0;
    #pragma unroll 1
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(mbar)), "r"(phase) : "memory");
    }
    phase ^= 1u;
    __syncwarp();
}

}  // namespace avk
