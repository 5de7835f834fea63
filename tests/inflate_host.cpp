// Host build of the BGZF inflate the device runs (aardvark_b200/csrc/avk_inflate.cuh): tests/test_bgzf.py compares it with zlib.
#include <cstring>
#include <vector>

#include "../aardvark_b200/csrc/avk_inflate.cuh"

using namespace avk_inflate;

extern "C" int inf_raw(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *out_len) {
    static thread_local Tables t;
    *out_len = 0;
    return inflate(in, n, out, cap, out_len, t);
}

// whole BGZF file -> out; returns 0, or 1000 * (member index + 1) + code (member walk errors: code 100 - rc)
extern "C" int inf_bgzf(const uint8_t *gz, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *out_len, int verify) {
    static thread_local Tables t;
    uint32_t tab[256];
    for (uint32_t i = 0; i < 256; ++i) tab[i] = crc_entry(i);
    uint64_t at = 0, o = 0;
    int k = 0;
    while (at < n) {
        Member m;
        uint64_t next = 0;
        const int rc = member_at(gz, n, at, m, next);
        if (rc != 0) return 1000 * (k + 1) + 100 - rc;
        if (o + m.isize > cap) return 1000 * (k + 1) + INF_E_OUTPUT;
        uint64_t got = 0;
        const int e = inflate(gz + m.c_off, m.c_len, out + o, m.isize, &got, t);
        if (e != INF_OK) return 1000 * (k + 1) + e;
        if (got != m.isize) return 1000 * (k + 1) + INF_E_SIZE;
        if (verify && crc32(tab, out + o, got) != m.crc) return 1000 * (k + 1) + INF_E_CRC;
        o += got; at = next; ++k;
    }
    *out_len = o;
    return 0;
}

// text -> BGZF file (members of DEFLATE_CHUNK bytes + the EOF marker) with the compressor the device runs; returns the length
extern "C" uint64_t def_bgzf(const uint8_t *text, uint64_t n, uint8_t *out, uint64_t cap) {
    static thread_local DeflateWork head;
    uint32_t t4[1024];
    crc_tables4(t4);
    uint64_t o = 0;
    std::vector<uint8_t> pay(DEFLATE_MAX_OUT + 8);
    for (uint64_t at = 0; at < n; at += DEFLATE_CHUNK) {
        const uint32_t len = (uint32_t)(n - at < DEFLATE_CHUNK ? n - at : DEFLATE_CHUNK);
        const uint32_t c = deflate_member(text + at, len, pay.data(), head);
        if (o + 18 + c + 8 > cap) return 0;
        member_header(out + o, c);
        memcpy(out + o + 18, pay.data(), c);
        member_trailer(out + o + 18 + c, crc32_4(t4, text + at, len), len);
        o += 18 + c + 8;
    }
    const uint32_t c = deflate_member(text, 0, pay.data(), head);      // EOF marker: an empty member
    if (o + 18 + c + 8 > cap) return 0;
    member_header(out + o, c);
    memcpy(out + o + 18, pay.data(), c);
    member_trailer(out + o + 18 + c, 0, 0);
    return o + 18 + c + 8;
}

// the length-limited Huffman construction on its own: lengths for freq[0..n) with at most `limit` bits
extern "C" void def_huff_lengths(const uint32_t *freq, int n, int limit, uint8_t *len) {
    static thread_local DeflateWork w;
    huff_lengths(w, freq, n, limit, len);
}
