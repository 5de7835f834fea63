// avk_vcf.cuh -- VCF body text -> call-set table on the device: parse_variant / parse_genotype /
// get_variant_type of src/parsing/region_generation.rs:565-758, one record line per thread.
//
// What stays on the host: BGZF inflate and the tabix query (noodles, src/parsing/noodles_helper.rs) -- the caller hands
// over the inflated record lines of whatever it queried -- and the header (sample names -> sample_index).  What moves here
// is the per-record work: column split, POS, the sample's GT (allele indices, phasing, '.' = reference, haploid =
// homozygous, :660-712), the split of a multi-ALT genotype into one variant per ALT allele (:594-653), '*' and symbolic
// ALTs dropped (:596-605), trailing-base trimming (:615-618), the 10 kbp allele limit (:621-626), type inference from
// INFO SVTYPE / TRID and the allele lengths (:719-758; BND and DUP records are dropped, :641-644), raw_allele_space (:610-612).
// Three passes: line starts (a select over the newline flags), k_vcf_scan (per line: how many variants and allele bytes it
// yields, or an error), exclusive scans, k_vcf_emit (writes the records where the scans say).
#pragma once
#include <stdint.h>

namespace avk_vcf {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

enum { VCF_OK = 0, VCF_E_COLUMNS = 1, VCF_E_POS = 2, VCF_E_NO_GT = 3, VCF_E_GT = 4, VCF_E_ALT_INDEX = 5, VCF_E_SVTYPE = 6, VCF_E_CONTIG = 7, VCF_E_EMPTY_ALLELE = 8, VCF_E_VARIANT = 9 };
enum { VCF_ALLELE_LIMIT = 10000 };

struct Field { u64 b, e; };       // [b, e) of the text

// per output variant of a line
struct Var { u32 alt_b, alt_e; u32 l0, l1, raw; u8 zyg, type, keep, pad; };
struct LineInfo { u32 pos; u32 contig; Field ref; int n; Var v[2]; int err; };

#if defined(__CUDACC__)
#define VCF_HD __host__ __device__
#else
#define VCF_HD
#endif

VCF_HD inline bool key_at(const u8 *t, u64 b, u64 e, const char *key, bool *has_value) {   // INFO field [b, e) is `key` or `key=...`
    u64 i = b;
    for (const char *k = key; *k; ++k, ++i) if (i >= e || t[i] != (u8)*k) return false;
    if (i == e) { *has_value = false; return true; }
    if (t[i] != '=') return false;
    *has_value = i + 1 < e;
    return true;
}

// One record line [b, e) (no trailing newline) -> LineInfo.  names: n_contigs zero-terminated strings of name_stride bytes.
VCF_HD inline void parse_line(const u8 *t, u64 b, u64 e, const char *names, u32 n_contigs, u32 name_stride, u32 sample_index, bool trim, LineInfo &L) {
    L.n = 0; L.err = VCF_OK; L.pos = 0; L.contig = 0;
    if (b >= e || t[b] == '#') return;                                           // header / empty line
    // columns: CHROM POS ID REF ALT QUAL FILTER INFO FORMAT sample...
    Field col[9];
    Field samp = {0, 0};
    int nc = 0;
    u64 s = b;
    for (u64 i = b; i <= e; ++i) {
        if (i == e || t[i] == '\t') {
            if (nc < 9) col[nc] = Field{s, i};
            else if ((u32)(nc - 9) == sample_index) samp = Field{s, i};
            nc += 1;
            s = i + 1;
        }
    }
    if (nc < 10 || (u32)(nc - 9) <= sample_index) { L.err = VCF_E_COLUMNS; return; }
    // CHROM
    u32 c = n_contigs;
    for (u32 k = 0; k < n_contigs && c == n_contigs; ++k) {
        const char *nm = names + (u64)k * name_stride;
        u64 i = col[0].b;
        bool same = true;
        for (u32 j = 0; nm[j] && same; ++j, ++i) same = i < col[0].e && t[i] == (u8)nm[j];
        if (same && i == col[0].e) c = k;
    }
    if (c == n_contigs) { L.err = VCF_E_CONTIG; return; }
    L.contig = c;
    // POS (1-based)
    u64 pos = 0;
    if (col[1].b == col[1].e) { L.err = VCF_E_POS; return; }
    for (u64 i = col[1].b; i < col[1].e; ++i) { if (t[i] < '0' || t[i] > '9') { L.err = VCF_E_POS; return; } pos = pos * 10 + (t[i] - '0'); if (pos > 0xffffffffull) { L.err = VCF_E_POS; return; } }
    if (pos == 0) { L.err = VCF_E_POS; return; }
    L.pos = (u32)(pos - 1);
    L.ref = col[3];
    // FORMAT: index of GT
    int gt_idx = -1, fi = 0;
    s = col[8].b;
    for (u64 i = col[8].b; i <= col[8].e; ++i)
        if (i == col[8].e || t[i] == ':') { if (i - s == 2 && t[s] == 'G' && t[s + 1] == 'T') gt_idx = fi; fi += 1; s = i + 1; }
    if (gt_idx < 0) { L.err = VCF_E_NO_GT; return; }
    // the sample's GT value
    Field gt = {samp.e, samp.e};
    fi = 0; s = samp.b;
    for (u64 i = samp.b; i <= samp.e; ++i)
        if (i == samp.e || t[i] == ':') { if (fi == gt_idx) gt = Field{s, i}; fi += 1; s = i + 1; }
    if (gt.b == gt.e) return;                                                     // trailing field dropped: missing value -> no-op (:582-587)
    if (gt.e - gt.b == 1 && t[gt.b] == '.') return;                               // GT = '.'
    // parse_genotype (:660-712): alleles separated by '/' or '|'; '.' = reference; one allele = homozygous
    int na = 0, a[2] = {0, 0};
    bool phased = false;
    u64 i = gt.b;
    if (i < gt.e && (t[i] == '/' || t[i] == '|')) ++i;                            // VCF 4.4 leading phasing marker
    for (;;) {
        if (na == 2) { L.err = VCF_E_GT; return; }
        if (i >= gt.e) { L.err = VCF_E_GT; return; }
        int val = 0;
        if (t[i] == '.') { ++i; }
        else {
            if (t[i] < '0' || t[i] > '9') { L.err = VCF_E_GT; return; }
            while (i < gt.e && t[i] >= '0' && t[i] <= '9') { val = val * 10 + (t[i] - '0'); if (val > 60000) { L.err = VCF_E_GT; return; } ++i; }
        }
        a[na++] = val;
        if (i == gt.e) break;
        if (t[i] == '|') phased = true; else if (t[i] != '/') { L.err = VCF_E_GT; return; }
        ++i;
    }
    if (na == 1) a[1] = a[0];
    int alt_idx[2], zyg[2], n = 0;
    if (a[0] == a[1]) { if (a[0] != 0) { alt_idx[0] = a[0]; zyg[0] = AVK_ZYG_HOM_ALT; n = 1; } }
    else {
        if (a[0] != 0) { alt_idx[n] = a[0]; zyg[n] = phased ? AVK_ZYG_PHASED_HET10 : AVK_ZYG_UNPHASED_HET; n += 1; }
        if (a[1] != 0) { alt_idx[n] = a[1]; zyg[n] = phased ? AVK_ZYG_PHASED_HET01 : AVK_ZYG_UNPHASED_HET; n += 1; }
    }
    if (n == 0) return;
    // INFO: SVTYPE / TRID (get_variant_type :723-746)
    int sv = -1;          // -1 none, else AVK_VT_SV_*
    bool trid = false, bad_sv = false;
    s = col[7].b;
    for (u64 j = col[7].b; j <= col[7].e; ++j)
        if (j == col[7].e || t[j] == ';') {
            bool hv = false;
            if (key_at(t, s, j, "SVTYPE", &hv) && hv) {
                const u64 vb = s + 7, vl = j - vb;
                if (vl == 3 && t[vb] == 'B' && t[vb + 1] == 'N' && t[vb + 2] == 'D') sv = AVK_VT_SV_BREAKEND;
                else if (vl == 3 && t[vb] == 'D' && t[vb + 1] == 'E' && t[vb + 2] == 'L') sv = AVK_VT_SV_DELETION;
                else if (vl == 3 && t[vb] == 'D' && t[vb + 1] == 'U' && t[vb + 2] == 'P') sv = AVK_VT_SV_DUPLICATION;
                else if (vl == 3 && t[vb] == 'I' && t[vb + 1] == 'N' && t[vb + 2] == 'S') sv = AVK_VT_SV_INSERTION;
                else bad_sv = true;
            }
            if (key_at(t, s, j, "TRID", &hv) && hv) trid = true;
            s = j + 1;
        }
    // ALT alleles
    for (int q = 0; q < n; ++q) {
        Field alt = {col[4].e, col[4].e};
        int ai = 1;
        s = col[4].b;
        for (u64 j = col[4].b; j <= col[4].e; ++j)
            if (j == col[4].e || t[j] == ',') { if (ai == alt_idx[q]) alt = Field{s, j}; ai += 1; s = j + 1; }
        if (alt_idx[q] >= ai) { L.err = VCF_E_ALT_INDEX; return; }                 // the reference would index out of bounds
        Var &V = L.v[L.n];
        V.keep = 0; V.zyg = (u8)zyg[q]; V.pad = 0;
        u64 l0 = L.ref.e - L.ref.b, l1 = alt.e - alt.b;
        if (l0 == 0 || l1 == 0) { L.err = VCF_E_EMPTY_ALLELE; return; }
        if (l1 == 1 && t[alt.b] == '*') continue;                                 // effectively a reference allele (:596-599)
        if (t[alt.b] == '<') continue;                                            // symbolic: needs sequence-resolved (:603-606)
        const u64 raw = l0 > l1 ? l0 : l1;                                        // before trimming (:610-612)
        if (trim) while (l0 > 1 && l1 > 1 && t[L.ref.b + l0 - 1] == t[alt.b + l1 - 1]) { --l0; --l1; }   // :615-618
        if (l0 > VCF_ALLELE_LIMIT || l1 > VCF_ALLELE_LIMIT) continue;             // :621-626
        int vt;
        if (bad_sv) { L.err = VCF_E_SVTYPE; return; }
        if (sv >= 0) vt = sv;
        else if (trid) vt = l1 < l0 ? AVK_VT_TR_CONTRACTION : AVK_VT_TR_EXPANSION;
        else vt = (l0 == 1 && l1 == 1) ? AVK_VT_SNV : (l0 == 1 ? AVK_VT_INSERTION : (l1 == 1 ? AVK_VT_DELETION : AVK_VT_INDEL));
        if (vt == AVK_VT_SV_BREAKEND || vt == AVK_VT_SV_DUPLICATION) continue;    // explicitly unsupported (:641-644)
        // the Variant constructors' checks (src/data_types/variants.rs:232-290): an SV deletion must not grow, an SV insertion must not shrink
        if ((vt == AVK_VT_SV_DELETION && (l0 <= 1 || l1 > l0)) || (vt == AVK_VT_SV_INSERTION && l1 < l0)) { L.err = VCF_E_VARIANT; return; }
        V.alt_b = (u32)(alt.b - b); V.alt_e = (u32)(alt.b - b + l1); V.l0 = (u32)l0; V.l1 = (u32)l1; V.raw = (u32)raw; V.type = (u8)vt; V.keep = 1;
        L.n += 1;
    }
}

}  // namespace avk_vcf
