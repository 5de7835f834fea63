// host_digest.h -- host restatement of k_prep_fill (avk_lib.cu): the cluster digest the device solvers read.  TEST HARNESS ONLY
// (shared by tests/ts_host.cpp and tests/sp_host.cpp).
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>

#include "../aardvark_b200/csrc/avk_layout.h"

namespace host_digest {
using namespace avk;

static uint32_t edit_distance(const uint8_t *a, uint32_t la, const uint8_t *b, uint32_t lb) {
    std::vector<uint32_t> prev(lb + 1), cur(lb + 1);
    for (uint32_t j = 0; j <= lb; ++j) prev[j] = j;
    for (uint32_t i = 1; i <= la; ++i) {
        cur[0] = i;
        for (uint32_t j = 1; j <= lb; ++j) cur[j] = std::min({prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (a[i - 1] != b[j - 1] ? 1u : 0u)});
        prev.swap(cur);
    }
    return prev[lb];
}

// k_prep_fill for one region: header + records in merged order + allele bytes
static void build_digest(const avk_region_batch *b, uint64_t r, std::vector<uint8_t> &dig) {
    const avk_variant_table &t = b->variants;
    const long long start = b->start[r], end = b->end[r];
    const uint64_t v0[2] = {b->var_off[r * 2], b->var_off[r * 2 + 1]};
    const int cnt[2] = {(int)(b->var_off[r * 2 + 1] - v0[0]), (int)(b->var_off[r * 2 + 2] - v0[1])};
    const int N = cnt[0] + cnt[1];
    bool invalid = false;
    long long s_l1 = 0, s_b0 = 0, s_al = 0, mx = start;
    for (int side = 0; side < 2; ++side)
        for (int i = 0; i < cnt[side]; ++i) {
            const uint64_t gv = v0[side] + i;
            const uint32_t l0 = t.a0_len[gv], l1 = t.a1_len[gv], p = t.position[gv];
            invalid = invalid || l0 == 0 || l1 == 0 || t.variant_type[gv] >= AVK_N_VARIANT_TYPES || t.zygosity[gv] > AVK_ZYG_HOM_ALT;
            invalid = invalid || (long long)p < start || (long long)p + l0 > end;
            if (i > 0) invalid = invalid || t.position[gv - 1] > p;
            s_l1 += l1; s_b0 += std::max(l0, l1); s_al += l0 + l1; mx = std::max<long long>(mx, (long long)p + l0);
        }
    dig.assign(PH_SIZE + (size_t)VI_SIZE * N + (size_t)s_al + 32, 0);
    int *hdr = (int *)dig.data();
    if (invalid) { hdr[PH_STATUS / 4] = AVK_ST_BAD_INPUT; return; }
    struct Ent { uint32_t pos; int side; uint64_t gv; };
    std::vector<Ent> order;
    for (int side = 0; side < 2; ++side) for (int i = 0; i < cnt[side]; ++i) order.push_back({t.position[v0[side] + i], side, v0[side] + (uint64_t)i});
    std::stable_sort(order.begin(), order.end(), [](const Ent &a, const Ent &c) { return a.pos < c.pos; });   // truth before query on ties
    uint32_t seen = 0;
    for (const Ent &e : order) seen |= 1u << t.variant_type[e.gv];
    uint8_t *recs = dig.data() + PH_SIZE, *alle = recs + (size_t)VI_SIZE * N;
    uint32_t acc = 0;
    for (int oi = 0; oi < N; ++oi) {
        const Ent &e = order[oi];
        uint32_t *rec = (uint32_t *)(recs + (size_t)VI_SIZE * oi);
        const uint32_t l0 = t.a0_len[e.gv], l1 = t.a1_len[e.gv];
        const uint8_t *src = t.allele_pool + t.allele_off[e.gv];
        rec[VI_POS / 4] = e.pos; rec[VI_L0 / 4] = l0; rec[VI_L1 / 4] = l1; rec[VI_AOFF / 4] = acc;
        rec[VI_ALTED / 4] = edit_distance(src, l0, src + l0, l1); rec[VI_RAW / 4] = t.raw_allele_space[e.gv]; rec[VI_GV / 4] = (uint32_t)e.gv;
        const uint32_t ty = t.variant_type[e.gv];
        rec[VI_FLAGS / 4] = ty | ((uint32_t)t.zygosity[e.gv] << 8) | ((e.side == 0 ? 1u : 0u) << 16) |
                            ((uint32_t)__builtin_popcount(seen & ((1u << ty) - 1)) << 24);
        memcpy(alle + acc, src, l0 + l1);
        acc += l0 + l1;
    }
    hdr[PH_STATUS / 4] = AVK_ST_OK; hdr[PH_N / 4] = N; hdr[PH_N0 / 4] = cnt[0]; hdr[PH_N1 / 4] = cnt[1];
    hdr[PH_SUM_L1 / 4] = (int)s_l1; hdr[PH_B0 / 4] = (int)s_b0; hdr[PH_SUM_ALLE / 4] = (int)s_al; hdr[PH_MAX_END / 4] = (int)mx;
    hdr[PH_NSLOTS / 4] = __builtin_popcount(seen);
    int k = 0;
    for (int ty = 0; ty < AVK_N_VARIANT_TYPES; ++ty) if (seen & (1u << ty)) dig[PH_SLOT_TYPE + (k++)] = (uint8_t)ty;
}

}  // namespace host_digest
