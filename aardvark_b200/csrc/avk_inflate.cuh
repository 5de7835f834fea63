// avk_inflate.cuh -- BGZF block inflate for the VCF ingest ("parallel block inflate").
//
// The reference reads its inputs through noodles' bgzf reader (src/parsing/noodles_helper.rs:14-97, region_generation.rs:489-541),
// i.e. through a DEFLATE library that is not part of the repository (noodles-bgzf -> flate2 / libdeflate, Cargo.lock).  What is
// restated here is therefore the published format: RFC 1951 (DEFLATE), RFC 1952 (gzip member, CRC-32) and the BGZF section of the
// SAM specification (gzip members of at most 64 KiB with a "BC" extra subfield holding the member size).  Parity is anchored on
// zlib: tests/test_bgzf.py compares this decoder, built for the host, with Python's zlib on every block type.
//
// BGZF members are independent: the host only walks the member headers (12 + XLEN bytes each) to find the payloads and the
// output offsets, and k_bgzf_inflate (avk_lib.cu) gives every member a warp -- lane 0 runs inflate() below into a 64 KiB window
// in shared memory, where its tables live too, and the warp writes the window out.  Huffman codes are decoded through a
// first-level table (9 bits for literals / lengths, 7 for distances) and canonically, bit by bit, beyond it.
// Host and device share this code (tests/inflate_host.cpp builds it with g++).
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifndef AVK_HD
#ifdef __CUDACC__
#define AVK_HD __host__ __device__
#else
#define AVK_HD
#endif
#endif

namespace avk_inflate {

enum {
    INF_OK = 0,
    INF_E_INPUT = 1,        // ran out of input
    INF_E_BTYPE = 2,        // block type 3
    INF_E_STORED = 3,       // LEN != ~NLEN
    INF_E_LENGTHS = 4,      // bad code-length data (repeat without a previous length, too many lengths, no end-of-block code)
    INF_E_CODE = 5,         // over-subscribed code, or an incomplete one that is not the single-code case
    INF_E_SYMBOL = 6,       // invalid literal/length or distance symbol
    INF_E_DISTANCE = 7,     // distance reaches before the start of the output
    INF_E_OUTPUT = 8,       // more output than the member's ISIZE
    INF_E_SIZE = 9,         // less output than ISIZE
    INF_E_CRC = 10          // CRC-32 of the output differs from the trailer's
};

struct Bits {
    const uint8_t *p;
    uint64_t n, pos;
    uint64_t buf;
    int cnt;
    AVK_HD void init(const uint8_t *src, uint64_t len) { p = src; n = len; pos = 0; buf = 0; cnt = 0; }
    // k <= 32.  A refill takes every byte that fits (the loads are independent of each other: one memory latency per ~7 bytes
    // instead of one per byte); past the end of the input it shifts in zero bits -- a peek may look past it, over() tells
    // whether bits that do not exist were CONSUMED.
    AVK_HD void need(int k) {
        if (cnt >= k) return;
        while (cnt <= 56) {
            const uint64_t b = pos < n ? p[pos] : 0;
            pos += 1;
            buf |= b << cnt;
            cnt += 8;
        }
    }
    AVK_HD uint32_t peek(int k) { need(k); return (uint32_t)(buf & ((1ull << k) - 1)); }
    AVK_HD void drop(int k) { buf >>= k; cnt -= k; }
    AVK_HD uint32_t get(int k) {
        if (k == 0) return 0;
        const uint32_t v = peek(k);
        drop(k);
        return v;
    }
    AVK_HD bool over() const { return pos * 8 - (uint64_t)cnt > n * 8; }      // more bits CONSUMED than the input holds
    AVK_HD void align() { const int r = cnt & 7; buf >>= r; cnt -= r; }      // stored block: skip to the byte boundary
};

// canonical Huffman code: count[l] codes of length l, symbols in code order
template <int NSYM>
struct Huff {
    uint16_t count[16];
    uint16_t symbol[NSYM];
    // > 0: incomplete, 0: complete, < 0: over-subscribed (RFC 1951 3.2.2)
    AVK_HD int build(const uint8_t *len, int n) {
        for (int l = 0; l < 16; ++l) count[l] = 0;
        for (int s = 0; s < n; ++s) count[len[s]] += 1;
        if (count[0] == n) return 0;                          // no codes at all: complete, but decoding anything fails
        int left = 1;
        for (int l = 1; l < 16; ++l) { left <<= 1; left -= count[l]; if (left < 0) return left; }
        uint16_t offs[16];
        offs[1] = 0;
        for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
        for (int s = 0; s < n; ++s) if (len[s]) symbol[offs[len[s]]++] = (uint16_t)s;
        return left;
    }
    AVK_HD int decode(Bits &b) const {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l < 16; ++l) {
            code |= (int)b.get(1);
            const int c = count[l];
            if (code - c < first) return symbol[index + (code - first)];
            index += c; first += c;
            first <<= 1; code <<= 1;
        }
        return -1;
    }
};

// First-level lookup: the next FAST bits of the stream (LSB first = the code's first bit in bit 0) -> (symbol << 4) | code length
// for every code of at most FAST bits; 0 = a longer code, decoded bit by bit.
template <int FAST>
struct FastTab {
    uint16_t e[1 << FAST];
    template <int NSYM>
    AVK_HD void build(const Huff<NSYM> &h) {
        for (int i = 0; i < (1 << FAST); ++i) e[i] = 0;
        int code = 0, index = 0;
        for (int l = 1; l <= FAST; ++l) {
            for (int k = 0; k < h.count[l]; ++k, ++code, ++index) {
                int rev = 0;
                for (int i = 0; i < l; ++i) rev |= ((code >> i) & 1) << (l - 1 - i);
                const uint16_t v = (uint16_t)((h.symbol[index] << 4) | l);
                for (int r = rev; r < (1 << FAST); r += 1 << l) e[r] = v;
            }
            code <<= 1;
        }
    }
    template <int NSYM>
    AVK_HD int decode(Bits &b, const Huff<NSYM> &h) const {
        const uint16_t v = e[b.peek(FAST)];
        if (v) { b.drop(v & 15); return v >> 4; }
        return h.decode(b);
    }
};

struct Tables {
    Huff<288> lit;
    Huff<30> dist;
    FastTab<9> flit;
    FastTab<7> fdist;
    uint8_t len[320];
};

// length of literal/length symbol 257..285 and its extra bits; distance of symbol 0..29 and its extra bits (RFC 1951 3.2.5)
AVK_HD inline int len_extra(int s) { return s < 265 || s == 285 ? 0 : (s - 261) >> 2; }
AVK_HD inline int len_base(int s) { return s < 265 ? s - 254 : s == 285 ? 258 : ((4 + ((s - 265) & 3)) << len_extra(s)) + 3; }
AVK_HD inline int dist_extra(int s) { return s < 4 ? 0 : (s >> 1) - 1; }
AVK_HD inline int dist_base(int s) { return s < 4 ? s + 1 : ((2 + (s & 1)) << dist_extra(s)) + 1; }

AVK_HD inline int fixed_tables(Tables &t) {
    for (int s = 0; s < 144; ++s) t.len[s] = 8;
    for (int s = 144; s < 256; ++s) t.len[s] = 9;
    for (int s = 256; s < 280; ++s) t.len[s] = 7;
    for (int s = 280; s < 288; ++s) t.len[s] = 8;
    t.lit.build(t.len, 288);
    for (int s = 0; s < 30; ++s) t.len[s] = 5;
    t.dist.build(t.len, 30);
    t.flit.build(t.lit); t.fdist.build(t.dist);
    return INF_OK;
}

AVK_HD inline int dynamic_tables(Bits &b, Tables &t) {
    const int nlen = (int)b.get(5) + 257, ndist = (int)b.get(5) + 1, ncode = (int)b.get(4) + 4;
    if (nlen > 286 || ndist > 30) return INF_E_LENGTHS;
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t cl[19];
    for (int i = 0; i < 19; ++i) cl[i] = 0;
    for (int i = 0; i < ncode; ++i) cl[order[i]] = (uint8_t)b.get(3);
    Huff<19> clh;
    if (clh.build(cl, 19) != 0) return INF_E_CODE;           // the code-length code must be complete
    int i = 0;
    while (i < nlen + ndist) {
        const int s = clh.decode(b);
        if (s < 0) return INF_E_SYMBOL;
        if (s < 16) { t.len[i++] = (uint8_t)s; continue; }
        int prev = 0, rep;
        if (s == 16) { if (i == 0) return INF_E_LENGTHS; prev = t.len[i - 1]; rep = 3 + (int)b.get(2); }
        else if (s == 17) rep = 3 + (int)b.get(3);
        else rep = 11 + (int)b.get(7);
        if (i + rep > nlen + ndist) return INF_E_LENGTHS;
        while (rep--) t.len[i++] = (uint8_t)prev;
    }
    if (b.over()) return INF_E_INPUT;
    if (t.len[256] == 0) return INF_E_LENGTHS;               // no end-of-block code
    int left = t.lit.build(t.len, nlen);
    if (left < 0 || (left > 0 && nlen - t.lit.count[0] != 1)) return INF_E_CODE;
    left = t.dist.build(t.len + nlen, ndist);
    if (left < 0 || (left > 0 && ndist - t.dist.count[0] != 1)) return INF_E_CODE;
    t.flit.build(t.lit); t.fdist.build(t.dist);
    return INF_OK;
}

// one raw DEFLATE stream -> out[0 .. *out_len); out_cap is the most it may produce
AVK_HD inline int inflate(const uint8_t *in, uint64_t in_len, uint8_t *out, uint64_t out_cap, uint64_t *out_len, Tables &t) {
    Bits b;
    b.init(in, in_len);
    uint64_t o = 0;
    int last;
    do {
        last = (int)b.get(1);
        const int type = (int)b.get(2);
        if (b.over()) return INF_E_INPUT;
        if (type == 0) {
            b.align();
            const uint32_t len = b.get(16), nlen = b.get(16);
            if (b.over()) return INF_E_INPUT;
            if ((len ^ 0xffffu) != nlen) return INF_E_STORED;
            if (o + len > out_cap) return INF_E_OUTPUT;
            // the bit buffer holds whole bytes now: drain it, then copy straight from the input
            uint32_t k = 0;
            while (k < len && b.cnt >= 8) { out[o++] = (uint8_t)b.get(8); ++k; }
            if (b.over()) return INF_E_INPUT;                  // (bytes past the end were drained)
            if (k < len) {                                     // the buffer is empty now: pos is the stream position
                if (b.pos + (len - k) > b.n) return INF_E_INPUT;
                for (; k < len; ++k) out[o++] = b.p[b.pos++];
            }
            continue;
        }
        if (type == 3) return INF_E_BTYPE;
        const int rc = type == 1 ? fixed_tables(t) : dynamic_tables(b, t);
        if (rc != INF_OK) return rc;
        for (;;) {
            int s = t.flit.decode(b, t.lit);
            if (s < 0) return b.over() ? INF_E_INPUT : INF_E_SYMBOL;
            if (b.over()) return INF_E_INPUT;
            if (s < 256) {
                if (o >= out_cap) return INF_E_OUTPUT;
                out[o++] = (uint8_t)s;
                continue;
            }
            if (s == 256) break;
            if (s > 285) return INF_E_SYMBOL;
            const int len = len_base(s) + (int)b.get(len_extra(s));
            const int ds = t.fdist.decode(b, t.dist);
            if (ds < 0 || ds > 29) return b.over() ? INF_E_INPUT : INF_E_SYMBOL;
            const uint64_t dist = (uint64_t)dist_base(ds) + b.get(dist_extra(ds));
            if (b.over()) return INF_E_INPUT;
            if (dist > o) return INF_E_DISTANCE;
            if (o + (uint64_t)len > out_cap) return INF_E_OUTPUT;
            for (int k = 0; k < len; ++k, ++o) out[o] = out[o - dist];      // byte by byte: the ranges may overlap (run-length case)
        }
    } while (!last);
    *out_len = o;
    return INF_OK;
}

// CRC-32 of RFC 1952 (reflected 0xEDB88320), table of 256 entries made by crc_table()
AVK_HD inline uint32_t crc_entry(uint32_t i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    return c;
}
AVK_HD inline uint32_t crc32(const uint32_t *tab, const uint8_t *p, uint64_t n) {
    uint32_t c = 0xffffffffu;
    for (uint64_t i = 0; i < n; ++i) c = tab[(c ^ p[i]) & 0xffu] ^ (c >> 8);
    return c ^ 0xffffffffu;
}

// slicing by 4: t4 = T0..T3 (1024 entries), T_k[i] = CRC of byte i followed by k zero bytes
AVK_HD inline void crc_tables4(uint32_t *t4) {
    for (uint32_t i = 0; i < 256; ++i) t4[i] = crc_entry(i);
    for (int t = 1; t < 4; ++t) for (uint32_t i = 0; i < 256; ++i) { const uint32_t p = t4[(t - 1) * 256 + i]; t4[t * 256 + i] = (p >> 8) ^ t4[p & 0xffu]; }
}
AVK_HD inline uint32_t crc32_4(const uint32_t *t4, const uint8_t *buf, uint32_t n) {
    uint32_t c = 0xffffffffu, i = 0;
    while (i < n && ((uintptr_t)(buf + i) & 3u)) { c = t4[(c ^ buf[i]) & 0xffu] ^ (c >> 8); ++i; }
    for (; i + 4 <= n; i += 4) {
        c ^= *(const uint32_t *)(buf + i);                                                  // (little endian)
        c = t4[768 + (c & 0xffu)] ^ t4[512 + ((c >> 8) & 0xffu)] ^ t4[256 + ((c >> 16) & 0xffu)] ^ t4[c >> 24];
    }
    for (; i < n; ++i) c = t4[(c ^ buf[i]) & 0xffu] ^ (c >> 8);
    return c ^ 0xffffffffu;
}

// ---- the other direction: one BGZF member's payload from at most 0xff00 bytes ------------------------------------------------
// (the reference compresses truth.vcf.gz / query.vcf.gz / passing.vcf.gz through noodles' multithreaded bgzf
// writer, src/writers/compare_parallel.rs:25-214, variant_merger.rs:124-147.  Any valid DEFLATE stream is a valid member: the
// bytes differ from the reference's compressor's, the inflated content does not.)  One final block over an LZ77 parse -- a hash
// of the next three bytes remembers the last two positions they were seen at; a short match gives way to a longer one that
// starts at the next byte.  The parse runs twice: once to count the
// symbols, from which length-limited Huffman codes are built (RFC 1951 3.2.7, dynamic block), once to emit; tiny inputs take
// the fixed codes (3.2.6), and a stored block replaces whatever does not pay, so a member never exceeds 64 KiB.
enum { DEFLATE_CHUNK = 0xff00, DEFLATE_HASH_BITS = 12, DEFLATE_MAX_OUT = DEFLATE_CHUNK + 5, DEFLATE_DYN_MIN = 128 };

struct BitWriter {
    uint8_t *p;
    uint32_t cap, pos;
    uint64_t buf;
    int cnt;
    bool full;
    AVK_HD void init(uint8_t *dst, uint32_t n) { p = dst; cap = n; pos = 0; buf = 0; cnt = 0; full = false; }
    AVK_HD void put(uint32_t v, int k) {                       // k <= 24 bits, LSB first
        buf |= (uint64_t)v << cnt; cnt += k;
        while (cnt >= 8) { if (pos < cap) p[pos] = (uint8_t)buf; else full = true; pos += 1; buf >>= 8; cnt -= 8; }
    }
    AVK_HD void finish() { if (cnt > 0) put(0, 8 - cnt); }
};
AVK_HD inline uint32_t bit_reverse(uint32_t code, int k) {     // Huffman codes go into the stream most-significant bit first
    uint32_t r = 0;
    for (int i = 0; i < k; ++i) r |= ((code >> i) & 1u) << (k - 1 - i);
    return r;
}
AVK_HD inline int msb_index(uint32_t v) { int m = 0; while (v >>= 1) ++m; return m; }
AVK_HD inline int len_symbol(int len) {
    if (len <= 10) return 254 + len;
    if (len == 258) return 285;
    const int l = len - 3, e = msb_index((uint32_t)l) - 2;
    return 261 + 4 * e + ((l >> e) & 3);
}
AVK_HD inline int dist_symbol(int dist) {
    const int d = dist - 1;
    if (d < 4) return d;
    const int m = msb_index((uint32_t)d);
    return 2 * m + ((d >> (m - 1)) & 1);
}

// everything deflate_member needs besides its input and output (shared memory on the device)
struct DeflateWork {
    uint16_t head[2 << DEFLATE_HASH_BITS];             // two most recent positions per hash bucket
    uint32_t lfreq[286], dfreq[30], cfreq[19];
    uint16_t lcode[286], dcode[30], ccode[19];         // bit-reversed canonical codes
    uint8_t llen[286], dlen[30], clen[19];
    uint8_t seq[316 + 4];                               // code lengths of both alphabets, in order
    uint32_t weight[2 * 286];                           // Huffman construction: leaves then internal nodes
    uint16_t parent[2 * 286];
};

// Code lengths of an optimal prefix code for freq[0..n), at most `limit` bits: the two lightest live nodes are merged until one
// is left (n <= 286: a quadratic scan is cheap); if the tree comes out deeper than the limit the weights are halved (rounding
// up, so used symbols stay used) and it is built again.  One used symbol gets one bit.
AVK_HD inline void huff_lengths(DeflateWork &w, const uint32_t *freq, int n, int limit, uint8_t *len) {
    int shift = 0;
    for (;;) {
        int live = 0;
        for (int i = 0; i < n; ++i) {
            w.weight[i] = freq[i] ? ((freq[i] - 1) >> shift) + 1 : 0;
            w.parent[i] = 0xffff;
            if (freq[i]) ++live;
        }
        if (live == 0) { for (int i = 0; i < n; ++i) len[i] = 0; return; }
        if (live == 1) { for (int i = 0; i < n; ++i) len[i] = freq[i] ? 1 : 0; return; }
        int nodes = n;
        while (live > 1) {
            int a = -1, b = -1;                          // lightest, second lightest among the nodes without a parent
            for (int i = 0; i < nodes; ++i) {
                if (w.parent[i] != 0xffff || w.weight[i] == 0) continue;
                if (a < 0 || w.weight[i] < w.weight[a]) { b = a; a = i; }
                else if (b < 0 || w.weight[i] < w.weight[b]) b = i;
            }
            w.weight[nodes] = w.weight[a] + w.weight[b];
            w.parent[nodes] = 0xffff;
            w.parent[a] = w.parent[b] = (uint16_t)nodes;
            nodes += 1; live -= 1;
        }
        int deepest = 0;
        for (int i = 0; i < n; ++i) {
            int d = 0;
            if (freq[i]) for (int j = i; w.parent[j] != 0xffff; j = w.parent[j]) ++d;
            len[i] = (uint8_t)d;
            if (d > deepest) deepest = d;
        }
        if (deepest <= limit) return;
        shift += 1;
    }
}
// canonical codes of RFC 1951 3.2.2 from the lengths, stored bit-reversed
AVK_HD inline void huff_codes(const uint8_t *len, int n, uint16_t *code) {
    int count[16], next[16];
    for (int l = 0; l < 16; ++l) count[l] = 0;
    for (int i = 0; i < n; ++i) count[len[i]] += 1;
    count[0] = 0;
    int c = 0;
    for (int l = 1; l < 16; ++l) { c = (c + count[l - 1]) << 1; next[l] = c; }
    for (int i = 0; i < n; ++i) code[i] = len[i] ? (uint16_t)bit_reverse((uint32_t)next[len[i]]++, len[i]) : 0;
}

struct CountSink {
    DeflateWork &w;
    AVK_HD bool stop() const { return false; }
    AVK_HD void lit(int b) { w.lfreq[b] += 1; }
    AVK_HD void match(int len, int dist) { w.lfreq[len_symbol(len)] += 1; w.dfreq[dist_symbol(dist)] += 1; }
};
struct EmitSink {                                       // with the codes in w (dynamic block)
    DeflateWork &w;
    BitWriter &o;
    AVK_HD bool stop() const { return o.full; }
    AVK_HD void lit(int b) { o.put(w.lcode[b], w.llen[b]); }
    AVK_HD void match(int len, int dist) {
        const int s = len_symbol(len), ds = dist_symbol(dist);
        o.put(w.lcode[s], w.llen[s]);
        o.put((uint32_t)(len - len_base(s)), len_extra(s));
        o.put(w.dcode[ds], w.dlen[ds]);
        o.put((uint32_t)(dist - dist_base(ds)), dist_extra(ds));
    }
};
AVK_HD inline void put_fixed_litlen(BitWriter &o, int s) {
    if (s < 144) o.put(bit_reverse(0x30 + s, 8), 8);
    else if (s < 256) o.put(bit_reverse(0x190 + s - 144, 9), 9);
    else if (s < 280) o.put(bit_reverse(s - 256, 7), 7);
    else o.put(bit_reverse(0xc0 + s - 280, 8), 8);
}
struct FixedSink {
    BitWriter &o;
    AVK_HD bool stop() const { return o.full; }
    AVK_HD void lit(int b) { put_fixed_litlen(o, b); }
    AVK_HD void match(int len, int dist) {
        const int s = len_symbol(len), ds = dist_symbol(dist);
        put_fixed_litlen(o, s);
        o.put((uint32_t)(len - len_base(s)), len_extra(s));
        o.put(bit_reverse((uint32_t)ds, 5), 5);
        o.put((uint32_t)(dist - dist_base(ds)), dist_extra(ds));
    }
};
AVK_HD inline uint32_t hash3(const uint8_t *p) { return (((uint32_t)p[0] << 16 | (uint32_t)p[1] << 8 | p[2]) * 2654435761u) >> (32 - DEFLATE_HASH_BITS); }
// Longest match for position i among the two most recent positions with the same hash (most recent first: on a tie the nearer
// one, whose distance costs fewer bits, wins); remembers i.
AVK_HD inline void lz_find(const uint8_t *in, uint32_t n, uint16_t *head, uint32_t i, int &len, int &dist) {
    len = 0; dist = 0;
    if (i + 3 > n) return;
    uint16_t *b = head + 2 * hash3(in + i);
    const int lim = (int)(n - i < 258u ? n - i : 258u);
    for (int k = 0; k < 2; ++k) {
        const uint32_t cand = b[k];
        if (cand == 0xffff || i - cand > 32768u) continue;
        int l = 0;
        while (l < lim && in[cand + l] == in[i + l]) ++l;
        if (l > len) { len = l; dist = (int)(i - cand); }
    }
    b[1] = b[0]; b[0] = (uint16_t)i;
}
AVK_HD inline void lz_remember(const uint8_t *in, uint32_t n, uint16_t *head, uint32_t j) {
    if (j + 3 > n) return;
    uint16_t *b = head + 2 * hash3(in + j);
    b[1] = b[0]; b[0] = (uint16_t)j;
}
// The LZ77 parse, greedy with one step of lazy evaluation (a short match gives way to a longer one starting at the next byte).
// It is deterministic, so the counting pass and the emitting pass see the same literals and matches.
template <class Sink>
AVK_HD inline void lz_parse(const uint8_t *in, uint32_t n, uint16_t *head, Sink &sink) {
    for (int i = 0; i < 2 * (1 << DEFLATE_HASH_BITS); ++i) head[i] = 0xffff;
    uint32_t i = 0;
    int len = 0, dist = 0;
    bool have = false;
    while (i < n && !sink.stop()) {
        if (!have) lz_find(in, n, head, i, len, dist);
        have = false;
        if (len < 3) { sink.lit(in[i]); i += 1; continue; }
        bool next_known = false;
        if (len < 32 && i + 1 < n) {
            int len2, dist2;
            lz_find(in, n, head, i + 1, len2, dist2);
            next_known = true;
            if (len2 > len) { sink.lit(in[i]); i += 1; len = len2; dist = dist2; have = true; continue; }
        }
        sink.match(len, dist);
        if (len <= 16) for (uint32_t j = i + (next_known ? 2 : 1); j < i + (uint32_t)len; ++j) lz_remember(in, n, head, j);   // positions inside short matches
        i += (uint32_t)len;
    }
}
// header of a dynamic block (3.2.7): HLIT, HDIST, HCLEN, the code-length code, then both alphabets' lengths run-length coded
AVK_HD inline void put_dynamic_header(DeflateWork &w, BitWriter &o) {
    int hlit = 286, hdist = 30;
    while (hlit > 257 && w.llen[hlit - 1] == 0) --hlit;
    while (hdist > 1 && w.dlen[hdist - 1] == 0) --hdist;
    const int total = hlit + hdist;
    for (int i = 0; i < hlit; ++i) w.seq[i] = w.llen[i];
    for (int i = 0; i < hdist; ++i) w.seq[hlit + i] = w.dlen[i];
    // two passes over the same run-length tokenisation: count, then emit
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 0) for (int i = 0; i < 19; ++i) w.cfreq[i] = 0;
        else {
            huff_lengths(w, w.cfreq, 19, 7, w.clen);
            huff_codes(w.clen, 19, w.ccode);
            const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            int hclen = 19;
            while (hclen > 4 && w.clen[order[hclen - 1]] == 0) --hclen;
            o.put((uint32_t)(hlit - 257), 5); o.put((uint32_t)(hdist - 1), 5); o.put((uint32_t)(hclen - 4), 4);
            for (int i = 0; i < hclen; ++i) o.put(w.clen[order[i]], 3);
        }
        int i = 0;
        while (i < total) {
            const int v = w.seq[i];
            int run = 1;
            while (i + run < total && w.seq[i + run] == v) ++run;
            int sym, extra = 0, ebits = 0, used;
            if (v == 0 && run >= 11) { used = run < 138 ? run : 138; sym = 18; extra = used - 11; ebits = 7; }
            else if (v == 0 && run >= 3) { used = run; sym = 17; extra = used - 3; ebits = 3; }
            else { used = 1; sym = v; }                                  // the length itself; repeats of a non-zero one follow as 16s
            if (pass == 0) w.cfreq[sym] += 1; else { o.put(w.ccode[sym], w.clen[sym]); o.put((uint32_t)extra, ebits); }
            i += used;
            if (v != 0 && sym == v) {                                     // repeats of a non-zero length: 16 copies the previous one 3..6 times
                int left = run - 1;
                while (left >= 3) {
                    const int r = left < 6 ? left : 6;
                    if (pass == 0) w.cfreq[16] += 1; else { o.put(w.ccode[16], w.clen[16]); o.put((uint32_t)(r - 3), 2); }
                    left -= r; i += r;
                }
            }
        }
    }
}
// Returns the payload length (<= n + 5).
AVK_HD inline uint32_t deflate_member(const uint8_t *in, uint32_t n, uint8_t *out, DeflateWork &w) {
    BitWriter o;
    o.init(out, n + 4);                                        // anything longer loses to a stored block
    if (n < DEFLATE_DYN_MIN) {
        o.put(1, 1); o.put(1, 2);                              // BFINAL = 1, BTYPE = 01
        FixedSink fs{o};
        lz_parse(in, n, w.head, fs);
        put_fixed_litlen(o, 256);
    } else {
        for (int i = 0; i < 286; ++i) w.lfreq[i] = 0;
        for (int i = 0; i < 30; ++i) w.dfreq[i] = 0;
        w.lfreq[256] = 1;
        CountSink cs{w};
        lz_parse(in, n, w.head, cs);
        huff_lengths(w, w.lfreq, 286, 15, w.llen);
        huff_lengths(w, w.dfreq, 30, 15, w.dlen);
        huff_codes(w.llen, 286, w.lcode);
        huff_codes(w.dlen, 30, w.dcode);
        o.put(1, 1); o.put(2, 2);                              // BFINAL = 1, BTYPE = 10
        put_dynamic_header(w, o);
        EmitSink es{w, o};
        lz_parse(in, n, w.head, es);
        o.put(w.lcode[256], w.llen[256]);
    }
    o.finish();
    if (!o.full && o.pos < n + 5) return o.pos;
    // stored: BFINAL = 1, BTYPE = 00, padding, LEN, NLEN, bytes
    out[0] = 1; out[1] = (uint8_t)n; out[2] = (uint8_t)(n >> 8); out[3] = (uint8_t)~n; out[4] = (uint8_t)(~n >> 8);
    for (uint32_t k = 0; k < n; ++k) out[5 + k] = in[k];
    return n + 5;
}
// the 18 bytes in front of a member's payload and the 8 behind it (SAM specification 4.1); bsize = total member size
AVK_HD inline void member_header(uint8_t *h, uint32_t payload_len) {
    const uint32_t bsize = 18 + payload_len + 8 - 1;
    const uint8_t fixed[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    for (int i = 0; i < 16; ++i) h[i] = fixed[i];
    h[16] = (uint8_t)bsize; h[17] = (uint8_t)(bsize >> 8);
}
AVK_HD inline void member_trailer(uint8_t *t, uint32_t crc, uint32_t isize) {
    for (int i = 0; i < 4; ++i) { t[i] = (uint8_t)(crc >> (8 * i)); t[4 + i] = (uint8_t)(isize >> (8 * i)); }
}

// ---- BGZF member walk (host): SAM specification 4.1 ---------------------------------------------------------------------
struct Member { uint64_t c_off; uint32_t c_len, isize, crc; uint64_t o_off; };
// returns 0 and the member at byte `at` (next = offset of the following member), or a negative code: -1 truncated, -2 not a
// gzip member with deflate + FEXTRA, -3 no BC subfield / inconsistent sizes
inline int member_at(const uint8_t *gz, uint64_t n, uint64_t at, Member &m, uint64_t &next) {
    if (at + 18 > n) return -1;
    const uint8_t *h = gz + at;
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return -2;
    const uint32_t xlen = h[10] | ((uint32_t)h[11] << 8);
    if (at + 12 + xlen > n) return -1;
    int64_t bsize = -1;
    for (uint32_t x = 0; x + 4 <= xlen;) {
        const uint8_t *f = h + 12 + x;
        const uint32_t slen = f[2] | ((uint32_t)f[3] << 8);
        if (f[0] == 'B' && f[1] == 'C' && slen == 2 && x + 6 <= xlen) bsize = (int64_t)(f[4] | ((uint32_t)f[5] << 8)) + 1;
        x += 4 + slen;
    }
    if (bsize < 0 || (uint64_t)bsize < 12ull + xlen + 8 || (h[3] & ~4)) return -3;     // (no name / comment / header CRC in BGZF members)
    if (at + (uint64_t)bsize > n) return -1;
    const uint8_t *tr = h + bsize - 8;
    m.c_off = at + 12 + xlen;
    m.c_len = (uint32_t)(bsize - 12 - xlen - 8);
    m.crc = tr[0] | ((uint32_t)tr[1] << 8) | ((uint32_t)tr[2] << 16) | ((uint32_t)tr[3] << 24);
    m.isize = tr[4] | ((uint32_t)tr[5] << 8) | ((uint32_t)tr[6] << 16) | ((uint32_t)tr[7] << 24);
    next = at + (uint64_t)bsize;
    return 0;
}

}  // namespace avk_inflate
