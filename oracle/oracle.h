/*
 * oracle.h -- entry points of the CPU ORACLE (oracle/aardvark_oracle.cpp).  TEST INFRASTRUCTURE ONLY.
 *
 * The oracle takes the product's batch / output structs (include/aardvark_b200.h) so that a batch can be
 * handed to either side and the outputs compared bit for bit.  Nothing under aardvark_b200/ includes
 * this header or links liboracle.so; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do (through oracle/oracle_py.py).
 */
#ifndef AARDVARK_ORACLE_H
#define AARDVARK_ORACLE_H

#include "../include/aardvark_b200.h"

#ifdef __cplusplus
extern "C" {
#endif


int orc_compare_batch(const avk_region_batch *batch, const uint8_t *const *contigs,
                      const uint64_t *contig_lens, uint32_t n_contigs,
                      const avk_compare_cfg *cfg, avk_compare_out *out,
                      int n_threads, avk_work_counters *work);
int orc_build_regions(const avk_callsets *in, uint64_t contig_len, uint32_t contig, uint32_t flank, uint64_t first_region_id,
                      avk_region_batch *out);
int orc_build_regions_bed(const avk_callsets *in, const uint32_t *variant_contig, const avk_bed_intervals *bed, const uint64_t *contig_lens,
                          uint32_t n_contigs, uint32_t flank, uint64_t first_region_id, avk_region_batch *out);
int orc_vcf_parse(const uint8_t *text, uint64_t len, const char *const *contig_names, uint32_t n_contigs, uint32_t sample_index, int enable_trimming,
                  avk_vcf_out *out, int32_t *err_code);
/* Stratifications::containments (stratifications.rs:108-118, 197-210) of every region's var_coordinates()
 * (compare_region.rs:63-74, incl. its last()-not-max end) as waffle_solver.rs:151-166 queries them: mask[r] bit s. */
int orc_containments(const avk_region_batch *batch, const avk_strat_intervals *strata, uint64_t *mask);
int orc_merge_batch(const avk_region_batch *batch, const uint8_t *const *contigs,
                    const uint64_t *contig_lens, uint32_t n_contigs,
                    const avk_merge_cfg *cfg, avk_merge_out *out,
                    int n_threads, avk_work_counters *work);


#ifdef __cplusplus
}
#endif
#endif /* AARDVARK_ORACLE_H */
