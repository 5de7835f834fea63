// avk_writers.h -- host side of the two writers that consume the compare path's results:
//   * avk_summary_write: SummaryWriter::write_summary (src/writers/summary.rs:166-221, write_group / write_category /
//     write_gt_category :243-420) over one GroupTypeMetrics table, i.e. the [AVK_N_GROUPS][AVK_N_METRICS] sums the kernels
//     accumulate (avk_compare_out::totals or one row of strat_totals);
//   * avk_vcf_records_write: VariantCategorizer::write_variants (src/writers/variant_categorizer.rs:178-237): the
//     GT:BD:EA:OA:RI record of every variant of one input, in region order.
// Plain C++17, no CUDA: the counters and labels were computed on the device, these functions only format them.
#pragma once
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/aardvark_b200.h"

namespace avk_writers {

// f64 -> text the way the csv crate serialises it (ryu::Buffer::format_finite -- shortest round-trip digits, then ryu's
// "pretty" layout: plain decimals while the decimal point lies within [-5, 16] digits of the first digit, a trailing ".0"
// for integers, otherwise d.ddde[-]x)
static inline std::string ryu_f64(double v) {
    if (v == 0.0) return "0.0";
    char sci[64];
    auto r = std::to_chars(sci, sci + sizeof(sci), v, std::chars_format::scientific);   // shortest round-trip, d[.ddd]e[+-]xx
    std::string s(sci, r.ptr);
    std::string out;
    size_t i = 0;
    if (s[0] == '-') { out = "-"; i = 1; }
    const size_t e = s.find('e');
    std::string digits;
    for (size_t k = i; k < e; ++k) if (s[k] != '.') digits.push_back(s[k]);
    const int exp10 = std::atoi(s.c_str() + e + 1);
    const int length = (int)digits.size();
    const int kk = exp10 + 1;                            // 10^(kk-1) <= |v| < 10^kk
    const int k = kk - length;                           // v = digits * 10^k
    if (0 <= k && kk <= 16) { out += digits; out.append((size_t)k, '0'); out += ".0"; }             // 1234e7 -> 12340000000.0
    else if (0 < kk && kk <= 16) { out += digits.substr(0, (size_t)kk); out += "."; out += digits.substr((size_t)kk); }   // 1234e-2 -> 12.34
    else if (-5 < kk && kk <= 0) { out += "0."; out.append((size_t)(-kk), '0'); out += digits; }     // 1234e-6 -> 0.001234
    else if (length == 1) { out += digits; out += "e"; out += std::to_string(kk - 1); }              // 1e30
    else { out += digits.substr(0, 1); out += "."; out += digits.substr(1); out += "e"; out += std::to_string(kk - 1); }   // 1234e30 -> 1.234e33
    return out;
}

static const char *const METRIC_NAMES[5] = {"GT", "HAP", "WEIGHTED_HAP", "BASEPAIR", "RECORD_BP"};          // grouped_metrics.rs:10-27 (serde renames)
static const int METRIC_COL[5] = {AVK_M_GT, AVK_M_HAP, AVK_M_WEIGHTED_HAP, AVK_M_BASEPAIR, AVK_M_RECORD_BP};
static const char *const TYPE_NAMES[AVK_N_VARIANT_TYPES] = {"Snv", "Insertion", "Deletion", "Indel", "SvInsertion", "SvDeletion", "SvDuplication",
                                                            "SvInversion", "SvBreakend", "TrContraction", "TrExpansion", "Unknown"};   // {variant_type:?}

struct Row4 { uint64_t tp, fn, qtp, qfp, fn_gt, fp_gt; };

static inline void emit_row(std::string &o, char d, const char *label, const char *comparison, const char *region_label, const char *variant_type,
                            const Row4 &m, bool gt) {
    // SummaryRow (summary.rs:33-114): compare_label comparison region_label filter variant_type truth_total truth_tp truth_fn
    // query_total query_tp query_fp metric_recall metric_precision metric_f1 truth_fn_gt query_fp_gt; None -> empty field
    o += label; o += d; o += comparison; o += d; o += region_label; o += d; o += "ALL"; o += d; o += variant_type; o += d;
    o += std::to_string(m.tp + m.fn); o += d; o += std::to_string(m.tp); o += d; o += std::to_string(m.fn); o += d;
    o += std::to_string(m.qtp + m.qfp); o += d; o += std::to_string(m.qtp); o += d; o += std::to_string(m.qfp); o += d;
    const bool hr = m.tp + m.fn > 0, hp = m.qtp + m.qfp > 0;                                        // summary_metrics.rs:48-74
    const double recall = hr ? (double)m.tp / (double)(m.tp + m.fn) : 0.0, precision = hp ? (double)m.qtp / (double)(m.qtp + m.qfp) : 0.0;
    if (hr) o += ryu_f64(recall);
    o += d;
    if (hp) o += ryu_f64(precision);
    o += d;
    if (hr && hp) {
        const double f1 = 2.0 * recall * precision / (recall + precision);
        if (f1 == f1) o += ryu_f64(f1); else o += "NaN";                                             // 0 / 0 when both are 0.0
    }
    o += d;
    if (gt) o += std::to_string(m.fn_gt);
    o += d;
    if (gt) o += std::to_string(m.fp_gt);
    o += '\n';
}

static inline Row4 row_of(const uint64_t *totals, int group, int metric) {
    const uint64_t *p = totals + (size_t)group * AVK_N_METRICS + METRIC_COL[metric];
    Row4 r = {p[0], p[1], p[2], p[3], 0, 0};
    if (metric == 0) { r.fn_gt = totals[(size_t)group * AVK_N_METRICS + AVK_M_GT_TRUTH_FN_GT]; r.fp_gt = totals[(size_t)group * AVK_N_METRICS + AVK_M_GT_QUERY_FP_GT]; }
    return r;
}
static inline bool empty4(const Row4 &r) { return r.tp + r.fn + r.qtp + r.qfp == 0; }                // SummaryMetrics::is_empty

// write_group (:243-274): for every requested metric the ALL row, the non-empty per-type rows in VariantType order (the
// BTreeMap's), then the non-empty joint rows JointIndel, JointStructuralVariant, JointTandemRepeat (:176-204)
static inline std::string summary_text(const uint64_t *totals, const uint8_t *metrics, uint32_t n_metrics, const char *compare_label,
                                       const char *region_label, char delim, bool header) {
    std::string o;
    if (header) {
        const char *cols[16] = {"compare_label", "comparison", "region_label", "filter", "variant_type", "truth_total", "truth_tp", "truth_fn", "query_total",
                                "query_tp", "query_fp", "metric_recall", "metric_precision", "metric_f1", "truth_fn_gt", "query_fp_gt"};
        for (int c = 0; c < 16; ++c) { o += cols[c]; o += c == 15 ? '\n' : delim; }
    }
    static const int joint_indel[] = {AVK_VT_INSERTION, AVK_VT_DELETION, AVK_VT_INDEL};
    static const int joint_sv[] = {AVK_VT_SV_INSERTION, AVK_VT_SV_DELETION, AVK_VT_SV_DUPLICATION, AVK_VT_SV_INVERSION, AVK_VT_SV_BREAKEND};
    static const int joint_tr[] = {AVK_VT_TR_EXPANSION, AVK_VT_TR_CONTRACTION};
    struct Joint { const char *label; const int *types; int n; };
    const Joint joints[3] = {{"JointIndel", joint_indel, 3}, {"JointStructuralVariant", joint_sv, 5}, {"JointTandemRepeat", joint_tr, 2}};
    for (uint32_t q = 0; q < n_metrics; ++q) {
        const int m = metrics[q];
        if (m < 0 || m > 4) continue;
        const bool gt = m == 0;
        emit_row(o, delim, compare_label, METRIC_NAMES[m], region_label, "ALL", row_of(totals, 0, m), gt);
        for (int t = 0; t < AVK_N_VARIANT_TYPES; ++t) {
            const Row4 r = row_of(totals, 1 + t, m);
            if (!empty4(r)) emit_row(o, delim, compare_label, METRIC_NAMES[m], region_label, TYPE_NAMES[t], r, gt);
        }
        for (const Joint &j : joints) {
            Row4 s = {0, 0, 0, 0, 0, 0};
            for (int k = 0; k < j.n; ++k) { const Row4 r = row_of(totals, 1 + j.types[k], m); s.tp += r.tp; s.fn += r.fn; s.qtp += r.qtp; s.qfp += r.qfp; s.fn_gt += r.fn_gt; s.fp_gt += r.fp_gt; }
            if (!empty4(s)) emit_row(o, delim, compare_label, METRIC_NAMES[m], region_label, j.label, s, gt);
        }
    }
    return o;
}

// write_variants (variant_categorizer.rs:178-237) for input `side` (0 truth, 1 query) of regions [lo, hi): one VCF body line per
// variant -- CHROM, 1-based POS, no ID / QUAL / FILTER / INFO, FORMAT GT:BD:EA:OA:RI and the sample column
static inline std::string vcf_records_text(const avk_region_batch *b, uint32_t side, const char *const *contig_names, const uint8_t *var_class,
                                           const uint8_t *var_expected, const uint8_t *var_observed, uint64_t lo, uint64_t hi) {
    static const char *const GT[6] = {".", "0/0", "0/1", "0|1", "1|0", "1/1"};                       // PhasedZygosity -> genotype string (:190-197)
    static const char *const BD[4] = {"UNK", "TP", "FN", "FP"};                                      // Classification (variant_metrics.rs:11-21)
    const avk_variant_table &t = b->variants;
    const uint64_t K = b->n_inputs;
    std::string o;
    for (uint64_t r = lo; r < hi; ++r)
        for (uint64_t v = b->var_off[r * K + side]; v < b->var_off[r * K + side + 1]; ++v) {
            o += contig_names[b->contig[r]]; o += '\t';
            o += std::to_string((uint64_t)t.position[v] + 1); o += "\t.\t";
            o.append((const char *)t.allele_pool + t.allele_off[v], t.a0_len[v]); o += '\t';
            o.append((const char *)t.allele_pool + t.allele_off[v] + t.a0_len[v], t.a1_len[v]);
            o += "\t.\t.\t.\tGT:BD:EA:OA:RI\t";
            o += GT[t.zygosity[v] <= AVK_ZYG_HOM_ALT ? t.zygosity[v] : 0]; o += ':';
            o += BD[var_class[v] <= AVK_CLASS_FP ? var_class[v] : 0]; o += ':';
            o += std::to_string((int)var_expected[v]); o += ':';
            o += std::to_string((int)var_observed[v]); o += ':';
            o += std::to_string((int32_t)b->region_id[r]);                                           // `region_id as i32` (:214)
            o += '\n';
        }
    return o;
}

}  // namespace avk_writers
