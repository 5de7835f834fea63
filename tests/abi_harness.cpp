// abi_harness.cpp -- a plain C++ caller of the C ABI (include/aardvark_b200.h), doing what the reference's run_compare does
// around its solve loop (src/main.rs:217-279): regions in, one batched call, per-region results and summary counters out.
// Built and run by tests/test_gpu_parity.py::test_cpp_abi_harness; no Python, no ctypes between this file and the library.
//
// The batch is the reference's first solve_compare_region test (waffle_solver.rs:897-930: one phased het SNV on both sides)
// plus a false-negative SNV, on mock_chr1.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../include/aardvark_b200.h"

int main() {
    const char *ref = "ACCGTTACCAGGACTTGACAAACCG";
    const uint8_t *seqs[1] = {(const uint8_t *)ref};
    const uint64_t lens[1] = {strlen(ref)};
    avk_ctx *ctx = nullptr;
    if (avk_create(0, &ctx) != AVK_OK) { fprintf(stderr, "avk_create failed\n"); return 2; }
    if (avk_set_reference(ctx, 1, seqs, lens) != AVK_OK) { fprintf(stderr, "%s\n", avk_last_error(ctx)); return 2; }

    // region 0: truth C>G at 2 (1|0), query the same; region 1: truth C>T at 13 (hom), query empty
    const uint64_t region_id[2] = {0, 1};
    const uint32_t contig[2] = {0, 0}, start[2] = {0, 10}, end[2] = {10, 20};
    const uint64_t var_off[5] = {0, 1, 2, 3, 3};
    const uint32_t pos[3] = {2, 2, 13};
    const uint8_t vtype[3] = {AVK_VT_SNV, AVK_VT_SNV, AVK_VT_SNV};
    const uint8_t zyg[3] = {AVK_ZYG_PHASED_HET10, AVK_ZYG_PHASED_HET10, AVK_ZYG_HOM_ALT};
    const uint32_t raw[3] = {1, 1, 1}, aoff[3] = {0, 2, 4}, l0[3] = {1, 1, 1}, l1[3] = {1, 1, 1};
    const uint8_t pool[6] = {'C', 'G', 'C', 'G', 'C', 'T'};
    avk_region_batch b;
    memset(&b, 0, sizeof(b));
    b.n_regions = 2; b.n_inputs = 2; b.region_id = region_id; b.contig = contig; b.start = start; b.end = end; b.var_off = var_off;
    b.variants.n_variants = 3; b.variants.position = pos; b.variants.variant_type = vtype; b.variants.zygosity = zyg;
    b.variants.raw_allele_space = raw; b.variants.allele_off = aoff; b.variants.a0_len = l0; b.variants.a1_len = l1;
    b.variants.allele_pool = pool; b.variants.allele_pool_len = sizeof(pool);

    int32_t status[2];
    uint32_t ed1[2], ed2[2];
    uint16_t type_mask[2], totals_mask = 0;
    uint8_t vexp[3], vobs[3], vcls[3];
    std::vector<uint64_t> totals(AVK_N_GROUPS * AVK_N_METRICS, 0);
    uint64_t solved = 0, errors = 0;
    avk_compare_out out;
    memset(&out, 0, sizeof(out));
    out.status = status; out.ed1 = ed1; out.ed2 = ed2; out.type_mask = type_mask;
    out.var_expected = vexp; out.var_observed = vobs; out.var_class = vcls;
    out.totals = totals.data(); out.totals_mask = &totals_mask; out.solved_blocks = &solved; out.error_blocks = &errors;
    const avk_compare_cfg cfg = {50, 0, 0, 0, 0};
    if (avk_compare_batch(ctx, &b, &cfg, &out) != AVK_OK) { fprintf(stderr, "%s\n", avk_last_error(ctx)); return 2; }
    printf("status %d %d ed %u %u %u %u solved %llu errors %llu\n", status[0], status[1], ed1[0], ed2[0], ed1[1], ed2[1],
           (unsigned long long)solved, (unsigned long long)errors);
    printf("variants");
    for (int v = 0; v < 3; ++v) printf(" %u/%u/%u", vexp[v], vobs[v], vcls[v]);
    printf("\njoint GT tp %llu fn %llu qtp %llu qfp %llu BASEPAIR tp %llu fn %llu qtp %llu qfp %llu\n",
           (unsigned long long)totals[AVK_M_GT], (unsigned long long)totals[AVK_M_GT + 1], (unsigned long long)totals[AVK_M_GT + 2],
           (unsigned long long)totals[AVK_M_GT + 3], (unsigned long long)totals[AVK_M_BASEPAIR], (unsigned long long)totals[AVK_M_BASEPAIR + 1],
           (unsigned long long)totals[AVK_M_BASEPAIR + 2], (unsigned long long)totals[AVK_M_BASEPAIR + 3]);
    avk_destroy(ctx);
    return 0;
}
