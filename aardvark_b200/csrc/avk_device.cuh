// avk_device.cuh -- warp-cooperative device primitives of the haplotype-comparison path.
//
// One warp owns one cluster (or one alignment).  All 32 lanes execute these
// functions together with warp-uniform scalar arguments; lanes split either
//   * the BASES of one diagonal (32 bytes per step, __ballot_sync + __ffs), or
//   * the DIAGONALS of one wavefront (one diagonal per lane),
// whichever the wavefront width calls for.  Nothing here uses tensor cores:
// this is integer DP, bounded by the INT32 ALU pipe.
//
// Semantics follow the reference exactly (paths relative to the reference repo):
//   DWFALite::{extend,increase_edit_distance,update,finalize}  src/dwfa/dynamic_wfa.rs:68-245
//   HaplotypeTracker / HaplotypeDWFA                            src/dwfa/haplotype_dwfa.rs:46-227
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define AVK_FULL 0xffffffffu

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;

namespace avk {

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- work counters (DESIGN.md "algorithmic work") ------------------------------------------
struct WorkAcc {
    unsigned long long alignments, cells, matched, search_pops, exact_pops;
    __device__ void clear() { alignments = cells = matched = search_pops = exact_pops = 0; }
};

// ---- byte copy ------------------------------------------------------------------------------
// dst/src may live in shared or global memory (generic pointers).
__device__ __forceinline__ void warp_copy(u8 *dst, const u8 *src, int n) {
    const int lane = lane_id();
    // head bytes until dst is 4-byte aligned, then words when src is equally aligned
    if (n >= 64 && (((uintptr_t)dst ^ (uintptr_t)src) & 3) == 0) {
        int head = (4 - ((uintptr_t)dst & 3)) & 3;
        if (lane < head) dst[lane] = src[lane];
        const u32 *s4 = (const u32 *)(src + head);
        u32 *d4 = (u32 *)(dst + head);
        int nw = (n - head) >> 2;
        for (int i = lane; i < nw; i += 32) d4[i] = s4[i];
        int done = head + (nw << 2);
        if (done + lane < n) dst[done + lane] = src[done + lane];
    } else {
        for (int i = lane; i < n; i += 32) dst[i] = src[i];
    }
}

// ---- longest common prefix, lanes across bases ----------------------------------------------
// number of equal bytes of a[0..na) and b[0..nb) from the start; warp-uniform result.
__device__ __forceinline__ int warp_lcp(const u8 *a, int na, const u8 *b, int nb) {
    const int lane = lane_id();
    const int maxn = min(na, nb);
    int total = 0;
    while (total < maxn) {
        int k = total + lane;
        bool eq = (k < maxn) && (a[k] == b[k]);
        unsigned m = __ballot_sync(AVK_FULL, eq);
        if (m == AVK_FULL) { total += 32; continue; }
        total += __ffs(~m) - 1;
        break;
    }
    return total;
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
// Wide wavefronts: one diagonal per lane; lanes that are still matching after a few
// bytes are finished with a warp-wide LCP so that one long run does not serialise the warp.
__device__ __forceinline__ Reach dwfa_extend(int *wf, int ed, const u8 *A, int la, const u8 *B, int lb, WorkAcc &w) {
    const int lane = lane_id();
    const int n = 2 * ed + 1;
    int mb = -1, mo = -1;
    bool full = false;
    int matched = 0;
    if (n <= 4) {
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
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const bool act = i < n;
            int d = act ? wf[i] : 0;
            int boff = d + ed - i;
            const int d0 = d;
            int cnt = 0;
            if (act) {
                while (cnt < 4 && boff < la && d < lb && A[boff] == B[d]) { ++d; ++boff; ++cnt; }
            }
            unsigned m = __ballot_sync(AVK_FULL, act && cnt == 4 && boff < la && d < lb);
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
    w.cells += (unsigned)n;
    w.matched += (unsigned)matched;
    Reach r;
    r.max_base = mb; r.max_other = mo; r.full = full;
    return r;
}

// increase_edit_distance() without the re-extend: dynamic_wfa.rs:152-168, in place.
// new[i] = max(old[i], old[i-1]+1, old[i-2]+1) over the entries that exist; chunks are
// processed from the top so that every read of old[] precedes the write that replaces it.
__device__ __forceinline__ void dwfa_grow(int *wf, int old_ed) {
    const int lane = lane_id();
    const int n_old = 2 * old_ed + 1;
    const int n_new = n_old + 2;
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
__device__ __forceinline__ int dwfa_update(int *wf, int *ed, int max_ed, const u8 *A, int la, const u8 *B, int lb, WorkAcc &w) {
    Reach r = dwfa_extend(wf, *ed, A, la, B, lb, w);
    while (!(r.max_base >= la) && !(r.max_other >= lb)) {
        *ed += 1;
        if (*ed > max_ed) return DWFA_MAX_ED;
        dwfa_grow(wf, *ed - 1);
        r = dwfa_extend(wf, *ed, A, la, B, lb, w);
    }
    return DWFA_OK;
}

// finalize(): dynamic_wfa.rs:183-198
__device__ __forceinline__ int dwfa_finalize(int *wf, int *ed, int max_ed, const u8 *A, int la, const u8 *B, int lb, WorkAcc &w) {
    w.alignments += 1;
    Reach r = dwfa_extend(wf, *ed, A, la, B, lb, w);
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
__device__ __forceinline__ int wfa_ed_warp(const u8 *A, int la, const u8 *B, int lb, int *wf, int max_ed, WorkAcc &w) {
    if (lane_id() == 0) wf[0] = 0;
    __syncwarp();
    int ed = 0;
    if (dwfa_finalize(wf, &ed, max_ed, A, la, B, lb, w) != DWFA_OK) return -1;
    return ed;
}

}  // namespace avk
