// sp_host.cpp -- HOST build of the speculative dense-cluster search (aardvark_b200/csrc/avk_spec_search.cuh).  TEST HARNESS ONLY.
//
// Chain and commit code are plain scalar C++; this file compiles the very same source for the CPU and runs the batch
// schedule with 32 simulated lanes (select_batch_host -> run_chain per lane -> commit_batch), so that the CPU test-suite can
// check the batched search against the oracle's optimize_sequences cluster by cluster (tests/test_spec_search_host.py)
// before it runs on a GPU.  The product never loads this library.
#include "../aardvark_b200/csrc/avk_spec_search.cuh"
#include "host_digest.h"

using namespace avk_sp;

// per region: status[r] = 0 searched, 1 rejected, 2 not tried (fewer than min_n variants); nres[r]; res[r][64][8] =
// a1 a2 ed1 ed2 tvs1 tvs2 qvs1 qvs2; tmask[r] = truth bits of the merged order
// stats: [0] clusters searched [1] rejected [2] batches [3] chains [4] pops [5] largest batch [6] sizeof(Shared) [7] sizeof(Scratch)
// scoring (when vexp != NULL): per variant expected / observed ALT copies of the chosen solution (query entries toggled as in
// compare_benchmark.rs:109-123), ed[r][2] = its two edit distances; status 3 = searched but the scoring was rejected
// stats: ... [8] exact-GT pops [9] exact-GT searches
extern "C" int sp_search_batch(const avk_region_batch *b, const uint8_t *const *contigs, const uint64_t *contig_lens, uint32_t n_contigs,
                               uint32_t mbf, uint32_t min_n, uint8_t *status, uint32_t *nres, uint32_t *res, uint32_t *tmask, uint64_t *stats,
                               uint8_t *vexp, uint8_t *vobs, uint32_t *ed) {
    if (!b || b->n_inputs != 2) return -1;
    std::vector<uint8_t> dig;
    Shared *S = new Shared();
    std::vector<Scratch> X(SP_LANES);
    Counters ctr = {0, 0, 0};
    for (int k = 0; k < 10; ++k) stats[k] = 0;
    for (uint64_t r = 0; r < b->n_regions; ++r) {
        status[r] = 2; nres[r] = 0;
        const uint64_t nv = b->var_off[r * 2 + 2] - b->var_off[r * 2];
        if (nv < min_n) continue;
        const uint32_t c = b->contig[r];
        status[r] = 1;
        if (c >= n_contigs || b->start[r] > b->end[r] || (uint64_t)b->end[r] > contig_lens[c] || b->end[r] > 0x7fff0000u) { stats[1] += 1; continue; }
        host_digest::build_digest(b, r, dig);
        if (!load_cluster(*S, dig.data(), (int)b->start[r], (int)b->end[r], (int)mbf)) { stats[1] += 1; continue; }
        View V;
        V.S = S; V.ref = contigs[c] + b->start[r]; V.recs = dig.data() + PH_SIZE; V.alle = V.recs + (size_t)VI_SIZE * S->N;
        bool ok = true;
        for (;;) {
            const int nb = select_batch_host(*S);
            if (nb == 0) break;
            stats[2] += 1; stats[3] += (uint64_t)nb; stats[5] = std::max<uint64_t>(stats[5], (uint64_t)nb);
            for (int l = 0; l < nb; ++l) { Group g; g.G = 1; g.sub = 0; g.mask = 1u << l; g.base = l; g.gs = &X[l]; run_chain(V, g, X[l], ctr, S->batch[l], S->out[l]); }
            if (!commit_batch(*S, nb)) { ok = false; break; }
        }
        if (!ok || S->nres == 0) { stats[1] += 1; continue; }
        status[r] = 0; nres[r] = (uint32_t)S->nres; tmask[r] = S->truth_mask;
        stats[0] += 1; stats[4] += S->spops;
        for (int i = 0; i < S->nres; ++i) {
            const ResEnt &e = S->res[i];
            uint32_t *o = res + (r * SP_RESCAP + (uint64_t)i) * 8;
            o[0] = e.a1; o[1] = e.a2; o[2] = e.ed1; o[3] = e.ed2; o[4] = e.tvs1; o[5] = e.tvs2; o[6] = e.qvs1; o[7] = e.qvs2;
        }
        if (!vexp) continue;
        score_prepare(*S, AVK_EXACT_GT_DEFAULT_MAX_EXPANSIONS);
        uint32_t xp = 0;
        for (int t = 0; t < S->n_tasks; ++t) {
            const ResEnt &e = S->res[S->task_r[t]];
            exact_gt_lane(V, X[t % SP_LANES], ctr, S->task_h[t] ? e.a2 : e.a1, 0x7fffffff, S->xcap, &xp, S->xout[t]);
        }
        stats[8] += xp; stats[9] += (uint64_t)S->n_tasks;
        if (!score_combine(*S)) { status[r] = 3; continue; }
        const ResEnt &best = S->res[S->best_r];
        ed[r * 2] = best.ed1; ed[r * 2 + 1] = best.ed2;
        const uint8_t *recs = dig.data() + PH_SIZE;
        for (int oi = 0; oi < S->N; ++oi) {
            const uint32_t *rec = (const uint32_t *)(recs + (size_t)VI_SIZE * oi);
            const int e_ = (int)((best.a1 >> oi) & 1) + (int)((best.a2 >> oi) & 1), o_ = (int)((S->keep0 >> oi) & 1) + (int)((S->keep1 >> oi) & 1);
            const bool tr = (S->truth_mask >> oi) & 1;
            vexp[rec[VI_GV / 4]] = (uint8_t)(tr ? e_ : o_); vobs[rec[VI_GV / 4]] = (uint8_t)(tr ? o_ : e_);
        }
    }
    stats[6] = sizeof(Shared); stats[7] = sizeof(Scratch);
    delete S;
    return 0;
}

// Whole solve of every region with at least min_n variants: search, scoring, final metrics, commit -- every output the
// product writes, into an avk_compare_out (rejected[r] = 1: outside the fast path, nothing written).
extern "C" int sp_solve_batch(const avk_region_batch *b, const uint8_t *const *contigs, const uint64_t *contig_lens, uint32_t n_contigs,
                              const avk_compare_cfg *cfg, uint32_t min_n, avk_compare_out *out, uint8_t *rejected, uint64_t *stats) {
    if (!b || b->n_inputs != 2 || !out || !rejected) return -1;
    std::vector<uint8_t> dig;
    Shared *S = new Shared();
    std::vector<Scratch> X(SP_LANES);
    Counters ctr = {0, 0, 0};
    for (int k = 0; k < 4; ++k) stats[k] = 0;
    for (uint64_t r = 0; r < b->n_regions; ++r) {
        rejected[r] = 1;
        if (cfg->enable_exact_shortcut || cfg->enable_sequences) continue;     // as k_search_spec: those clusters go to the warp solver
        const uint64_t nv = b->var_off[r * 2 + 2] - b->var_off[r * 2];
        if (nv < min_n) continue;
        stats[0] += 1;
        const uint32_t c = b->contig[r];
        if (c >= n_contigs || b->start[r] > b->end[r] || (uint64_t)b->end[r] > contig_lens[c] || b->end[r] > 0x7fff0000u) continue;
        host_digest::build_digest(b, r, dig);
        if (!load_cluster(*S, dig.data(), (int)b->start[r], (int)b->end[r], (int)cfg->max_branch_factor)) continue;
        View V;
        V.S = S; V.ref = contigs[c] + b->start[r]; V.recs = dig.data() + PH_SIZE; V.alle = V.recs + (size_t)VI_SIZE * S->N;
        bool ok = true;
        for (;;) {
            const int nb = select_batch_host(*S);
            if (nb == 0) break;
            for (int l = 0; l < nb; ++l) { Group g; g.G = 1; g.sub = 0; g.mask = 1u << l; g.base = l; g.gs = &X[l]; run_chain(V, g, X[l], ctr, S->batch[l], S->out[l]); }
            if (!commit_batch(*S, nb)) { ok = false; break; }
        }
        if (!ok || S->nres == 0) continue;
        score_prepare(*S, cfg->exact_gt_max_expansions ? cfg->exact_gt_max_expansions : AVK_EXACT_GT_DEFAULT_MAX_EXPANSIONS);
        uint32_t xp = 0;
        for (int t = 0; t < S->n_tasks; ++t) {
            const ResEnt &e = S->res[S->task_r[t]];
            exact_gt_lane(V, X[t % SP_LANES], ctr, S->task_h[t] ? e.a2 : e.a1, 0x7fffffff, S->xcap, &xp, S->xout[t]);
        }
        if (!score_combine(*S)) continue;
        if (S->n_slots > SP_MAXSLOT || !metrics_walk(V, *S, true)) continue;
        stats[2] += (uint64_t)S->n_mtasks;
        for (int t = 0; t < S->n_mtasks; ++t) metrics_align(V, X[t % SP_LANES], ctr, S->mtask[t]);
        if (!metrics_walk(V, *S, false)) continue;
        rejected[r] = 0;
        stats[1] += 1;
        const uint64_t v0 = b->var_off[r * 2], v1 = b->var_off[r * 2 + 2];
        uint64_t *row = out->region_metrics ? out->region_metrics + r * (uint64_t)(AVK_N_GROUPS * AVK_N_METRICS) : nullptr;
        if (row) memset(row, 0, sizeof(uint64_t) * AVK_N_GROUPS * AVK_N_METRICS);
        struct Sink {
            const View &V; avk_compare_out *out; uint64_t *row;
            void variant(int oi, int e, int o) {
                const uint32_t *rec = (const uint32_t *)(V.recs + (size_t)VI_SIZE * oi);
                const uint32_t gv = rec[VI_GV / 4];
                const bool tr = (rec[VI_FLAGS / 4] & 0x10000u) != 0;
                out->var_expected[gv] = (uint8_t)e; out->var_observed[gv] = (uint8_t)o;
                out->var_class[gv] = e == o ? AVK_CLASS_TP : (tr ? AVK_CLASS_FN : AVK_CLASS_FP);
            }
            void metric(int g, int m, uint64_t v) { if (row) row[g * AVK_N_METRICS + m] = v; }
        } sink{V, out, row};
        uint32_t e1 = 0, e2 = 0; uint16_t tm = 0;
        const int rc = commit_metrics(V, *S, sink, &e1, &e2, &tm);
        out->status[r] = rc;
        if (rc == AVK_ST_OK) { out->ed1[r] = e1; out->ed2[r] = e2; out->type_mask[r] = tm; }
        else {
            if (row) memset(row, 0, sizeof(uint64_t) * AVK_N_GROUPS * AVK_N_METRICS);
            out->ed1[r] = 0; out->ed2[r] = 0; out->type_mask[r] = 0;
            for (uint64_t v = v0; v < v1; ++v) { out->var_expected[v] = 0; out->var_observed[v] = 0; out->var_class[v] = AVK_CLASS_UNKNOWN; }
        }
    }
    delete S;
    return 0;
}
