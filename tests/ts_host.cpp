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
#include "host_digest.h"

using namespace avk_ts;

using host_digest::build_digest;

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
