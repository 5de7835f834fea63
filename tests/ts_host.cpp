// ts_host.cpp -- HOST build of the thread-per-cluster solver (aardvark_b200/csrc/avk_thread_solver.cuh).  TEST HARNESS ONLY.
//
// The thread solver is plain scalar C++; this file compiles the very same source for the CPU so that the CPU test-suite
// can check it against the oracle cluster by cluster (tests/test_thread_solver_host.py) before it runs on a GPU.  The
// product never loads this library: libaardvark_b200.so runs the solver in a CUDA kernel (k_compare_thread) and has no
// CPU path.  The digest builder below restates k_prep_fill (avk_lib.cu) for the host.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../aardvark_b200/csrc/avk_thread_solver.cuh"

using namespace avk_ts;

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

// stats: [0] accepted, [1] rejected, [2] search pops, [3] exact pops, [4] cells, [5] sizeof(Work)
extern "C" int ts_compare_batch(const avk_region_batch *b, const uint8_t *const *contigs, const uint64_t *contig_lens, uint32_t n_contigs,
                                const avk_compare_cfg *cfg, avk_compare_out *out, uint8_t *rejected, uint64_t *stats) {
    if (!b || b->n_inputs != 2 || !out || !rejected) return -1;
    std::vector<uint8_t> dig;
    Counters ctr = {0, 0, 0, 0, 0};
    uint64_t steps = 0;
    uint64_t acc = 0, rej = 0;
    Work *w = new Work();
    for (uint64_t r = 0; r < b->n_regions; ++r) {
        rejected[r] = 1;
        const uint32_t c = b->contig[r];
        if (c >= n_contigs || b->start[r] > b->end[r] || (uint64_t)b->end[r] > contig_lens[c] || b->end[r] > 0x7fff0000u) { rej += 1; continue; }
        if (cfg->enable_exact_shortcut || cfg->enable_sequences) { rej += 1; continue; }
        build_digest(b, r, dig);
        Solver S;
        S.wp = w; S.ctr = &ctr;
        S.begin(dig.data(), contigs[c], (int)b->start[r], (int)b->end[r], (int)cfg->max_branch_factor, cfg->exact_gt_max_expansions);
        while (S.phase == PH_RUN) {                                      // the kernel alternates these two for a whole warp
            S.advance();
            if (S.task.kind != TK_NONE) { S.exec_task(); steps += 1; }
        }
        int rc = S.rc;
        const Cluster &cl = S.c;
        if (rc == TS_REJECT) { rej += 1; continue; }
        const uint64_t v0 = b->var_off[r * 2], v1 = b->var_off[r * 2 + 2];
        uint64_t *row = out->region_metrics ? out->region_metrics + r * (uint64_t)(AVK_N_GROUPS * AVK_N_METRICS) : nullptr;
        if (row) memset(row, 0, sizeof(uint64_t) * AVK_N_GROUPS * AVK_N_METRICS);
        if (rc == AVK_ST_OK) {
            struct Sink {
                const Cluster &cl; avk_compare_out *out; uint64_t *row;
                void variant(int oi, int e, int o) {
                    const uint32_t gv = rec32(cl, oi, VI_GV);
                    const bool tr = (rec32(cl, oi, VI_FLAGS) & 0x10000u) != 0;
                    out->var_expected[gv] = (uint8_t)e; out->var_observed[gv] = (uint8_t)o;
                    out->var_class[gv] = e == o ? AVK_CLASS_TP : (tr ? AVK_CLASS_FN : AVK_CLASS_FP);
                }
                void metric(int g, int m, uint64_t v) { if (row) row[g * AVK_N_METRICS + m] = v; }
            } sink{cl, out, row};
            uint32_t e1 = 0, e2 = 0; uint16_t tm = 0;
            rc = commit_solution(S, sink, &e1, &e2, &tm);
            if (rc == AVK_ST_OK) { out->ed1[r] = e1; out->ed2[r] = e2; out->type_mask[r] = tm; }
        }
        rejected[r] = 0;
        acc += 1;
        out->status[r] = rc;
        if (rc != AVK_ST_OK) {
            if (row) memset(row, 0, sizeof(uint64_t) * AVK_N_GROUPS * AVK_N_METRICS);
            out->ed1[r] = 0; out->ed2[r] = 0; out->type_mask[r] = 0;
            for (uint64_t v = v0; v < v1; ++v) { out->var_expected[v] = 0; out->var_observed[v] = 0; out->var_class[v] = AVK_CLASS_UNKNOWN; }
        }
    }
    delete w;
    if (stats) { stats[0] = acc; stats[1] = rej; stats[2] = ctr.spops; stats[3] = ctr.xpops; stats[4] = ctr.cells; stats[5] = sizeof(Work); stats[6] = steps; }
    return 0;
}
