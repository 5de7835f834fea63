"""BGZF inflate.  The reference inflates through noodles-bgzf (flate2 / libdeflate; not in the repository), so
parity is anchored on zlib: the decoder the device runs (aardvark_b200/csrc/avk_inflate.cuh), built here for the host by
tests/inflate_host.cpp, must reproduce what Python's zlib produces -- stored, fixed-Huffman and dynamic-Huffman blocks, streams
of several blocks, the degenerate codes (one distance code, no distance codes), empty members -- and must reject damaged input."""
import ctypes as C
import os
import random
import struct
import subprocess
import zlib

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libinflate_host.so")
SRCS = [os.path.join(HERE, "inflate_host.cpp"), os.path.join(HERE, "..", "aardvark_b200", "csrc", "avk_inflate.cuh")]


def inf_lib():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(s) for s in SRCS):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", SO, SRCS[0]])
    lib = C.CDLL(SO)
    lib.inf_raw.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.inf_bgzf.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int]
    lib.def_bgzf.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]
    lib.def_bgzf.restype = C.c_uint64
    lib.def_huff_lengths.argtypes = [C.POINTER(C.c_uint32), C.c_int, C.c_int, C.POINTER(C.c_uint8)]
    return lib


def raw_inflate(data: bytes, cap: int):
    out = C.create_string_buffer(max(cap, 1))
    n = C.c_uint64(0)
    rc = inf_lib().inf_raw(data, len(data), out, cap, C.byref(n))
    return rc, out.raw[:n.value]


def deflate_raw(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


def bgzf_member(chunk: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, extra=b"") -> bytes:
    """One BGZF member (SAM spec 4.1); `extra`: further subfields in front of BC (allowed by the format)."""
    cdata = deflate_raw(chunk, level, strategy)
    xlen = 6 + len(extra)
    bsize = 12 + xlen + len(cdata) + 8 - 1
    assert bsize < 65536
    return (b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", xlen) + extra + b"BC" + struct.pack("<HH", 2, bsize) + cdata
            + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))


BGZF_EOF = bgzf_member(b"")


def bgzf_compress(data: bytes, level=6, block=0xff00, strategy=zlib.Z_DEFAULT_STRATEGY, eof=True) -> bytes:
    out = b"".join(bgzf_member(data[i:i + block], level, strategy) for i in range(0, len(data), block))
    return out + (BGZF_EOF if eof else b"")


def bgzf_inflate_host(gz: bytes, cap: int, verify=1):
    out = C.create_string_buffer(max(cap, 1))
    n = C.c_uint64(0)
    rc = inf_lib().inf_bgzf(gz, len(gz), out, cap, C.byref(n), verify)
    return rc, out.raw[:n.value]


def _texts():
    rnd = random.Random(7)
    vcf = b"".join(b"chr%d\t%d\t.\t%s\t%s\t.\tPASS\tSVTYPE=INS\tGT:AD\t0|1:%d,%d\n" % (rnd.randint(1, 22), rnd.randint(1, 10**8), rnd.choice([b"A", b"ACGT", b"G"]),
                   rnd.choice([b"T", b"TTTTTTTTTTTTTTTTTTTTTTTTTTTTTT", b"C"]), rnd.randint(0, 60), rnd.randint(0, 60)) for _ in range(3000))
    return {
        "empty": b"",
        "one": b"A",
        "run": b"A" * 70000,                                      # one distance code, matches overlapping their source
        "two_symbols": b"ABABABABAB" * 500,
        "text": vcf,
        "random": bytes(rnd.getrandbits(8) for _ in range(40000)),   # incompressible: zlib stores it
        "acgt": bytes(rnd.choice(b"ACGT") for _ in range(100000)),   # literals only in practice: few or no distance codes
        "far": bytes(rnd.getrandbits(8) for _ in range(32768)) * 2 + b"tail",   # matches at the maximum distance
    }


@pytest.mark.parametrize("name", sorted(_texts()))
@pytest.mark.parametrize("level,strategy", [(0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                                            (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)])
def test_raw_deflate_streams_vs_zlib(name, level, strategy):
    data = _texts()[name]
    comp = deflate_raw(data, level, strategy)
    assert zlib.decompress(comp, -15) == data
    rc, out = raw_inflate(comp, len(data))
    assert rc == 0 and out == data


def test_stream_of_many_blocks_and_flush_points():
    """Z_FULL_FLUSH / Z_SYNC_FLUSH insert empty stored blocks and restart the codes: several blocks of every type in one stream."""
    rnd = random.Random(3)
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    parts, data = [], b""
    for i in range(40):
        chunk = bytes(rnd.choice(b"ACGT\n\t01|/") for _ in range(rnd.randint(0, 3000))) if i % 3 else bytes(rnd.getrandbits(8) for _ in range(rnd.randint(0, 500)))
        data += chunk
        parts.append(c.compress(chunk) + c.flush(zlib.Z_FULL_FLUSH if i % 2 else zlib.Z_SYNC_FLUSH))
    comp = b"".join(parts) + c.flush()
    rc, out = raw_inflate(comp, len(data))
    assert rc == 0 and out == data


@settings(max_examples=150, deadline=None)
@given(st.binary(max_size=5000), st.sampled_from([0, 1, 6, 9]), st.sampled_from([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED]))
def test_raw_deflate_property(data, level, strategy):
    rc, out = raw_inflate(deflate_raw(data, level, strategy), len(data))
    assert rc == 0 and out == data


@settings(max_examples=60, deadline=None)
@given(st.lists(st.sampled_from([b"A", b"C", b"G", b"T", b"\t", b"\n", b"0|1", b"chr1", b"PASS"]), max_size=4000), st.integers(1, 5000))
def test_bgzf_property(tokens, block):
    data = b"".join(tokens)
    rc, out = bgzf_inflate_host(bgzf_compress(data, 6, block), len(data))
    assert rc == 0 and out == data


def test_bgzf_file_layout():
    data = _texts()["text"]
    gz = bgzf_compress(data, 6)
    assert zlib.decompress(gz, 31) == data[:0xff00]                 # (each member is a gzip member)
    rc, out = bgzf_inflate_host(gz, len(data))
    assert rc == 0 and out == data
    # extra subfields before BC, no EOF marker, members of one byte
    gz2 = bgzf_member(b"hello ", extra=b"XY" + struct.pack("<H", 3) + b"abc") + bgzf_member(b"") + bgzf_member(b"w") + bgzf_member(b"orld")
    rc, out = bgzf_inflate_host(gz2, 64)
    assert rc == 0 and out == b"hello world"
    assert bgzf_inflate_host(b"", 8) == (0, b"")


def test_damaged_input_is_rejected():
    data = _texts()["text"][:20000]
    gz = bytearray(bgzf_compress(data, 6))
    rc, _ = bgzf_inflate_host(bytes(gz[:-40]), len(data))            # truncated inside the EOF member
    assert rc != 0
    bad = bytearray(gz); bad[len(gz) // 2] ^= 0x55                   # a flipped payload byte: some decode error or the CRC
    rc, _ = bgzf_inflate_host(bytes(bad), len(data))
    assert rc != 0
    plain_gzip = zlib.compressobj(6, zlib.DEFLATED, 31)
    g = plain_gzip.compress(data) + plain_gzip.flush()
    rc, _ = bgzf_inflate_host(g, len(data))                          # a gzip file that is not BGZF (no FEXTRA / BC)
    assert rc % 1000 in (102, 103)
    # raw streams: block type 3, stored length mismatch, distance before the start, input that ends early
    assert raw_inflate(bytes([0b111]), 16)[0] == 2
    assert raw_inflate(bytes([0b001, 5, 0, 0, 0]), 16)[0] == 3
    comp = deflate_raw(b"abcabcabcabc" * 10, 6, zlib.Z_FIXED)
    assert raw_inflate(comp[:len(comp) // 2], 200)[0] == 1
    assert raw_inflate(comp, 10)[0] == 8                              # more output than the caller allows
    # fixed block: literal 'a', then length 3 at distance 2 with one byte of history
    stream = _fixed_stream([("lit", ord("a")), ("match", 3, 2), ("end",)])
    assert raw_inflate(stream, 16)[0] == 7


def _fixed_stream(ops):
    """Assemble a final fixed-Huffman block from literals / matches (RFC 1951 3.2.6) -- Huffman codes MSB first, extra bits LSB first."""
    bits = []
    def put_code(code, n):
        bits.extend((code >> (n - 1 - i)) & 1 for i in range(n))
    def put_extra(v, n):
        bits.extend((v >> i) & 1 for i in range(n))
    def lit(s):
        if s < 144: put_code(0x30 + s, 8)
        elif s < 256: put_code(0x190 + s - 144, 9)
        elif s < 280: put_code(s - 256, 7)
        else: put_code(0xc0 + s - 280, 8)
    put_extra(1, 1); put_extra(1, 2)                                 # BFINAL = 1, BTYPE = 01
    for op in ops:
        if op[0] == "lit":
            lit(op[1])
        elif op[0] == "end":
            lit(256)
        else:
            _, length, dist = op
            assert 3 <= length <= 10 and 1 <= dist <= 4              # symbols without extra bits
            lit(254 + length)
            put_code(dist - 1, 5)
    while len(bits) % 8:
        bits.append(0)
    return bytes(sum(b << i for i, b in enumerate(bits[k:k + 8])) for k in range(0, len(bits), 8))


def test_hand_assembled_fixed_block():
    s = _fixed_stream([("lit", ord("a")), ("lit", ord("b")), ("match", 4, 2), ("match", 3, 1), ("end",)])
    assert zlib.decompress(s, -15) == b"abababbbb"
    assert raw_inflate(s, 16) == (0, b"abababbbb")


# ---- compression: any valid BGZF will do; what is pinned is that zlib (and our own inflate) give the text back ------------------

def gunzip_members(gz: bytes) -> bytes:
    out, buf = [], gz
    while buf:
        d = zlib.decompressobj(31)
        out.append(d.decompress(buf))
        assert d.eof
        buf = d.unused_data
    return b"".join(out)


def bgzf_compress_host(text: bytes) -> bytes:
    cap = len(text) + (len(text) // 0xff00 + 2) * 64
    out = C.create_string_buffer(cap)
    n = inf_lib().def_bgzf(text, len(text), out, cap)
    assert n > 0
    return out.raw[:n]


def check_bgzf_layout(gz: bytes, text: bytes):
    """members of at most 64 KiB with the BC subfield, the sizes and CRCs of the chunks, the 28-byte EOF marker at the end"""
    at, chunks = 0, []
    while at < len(gz):
        assert gz[at:at + 4] == b"\x1f\x8b\x08\x04" and gz[at + 10:at + 16] == b"\x06\x00BC\x02\x00"
        bsize = struct.unpack("<H", gz[at + 16:at + 18])[0] + 1
        crc, isize = struct.unpack("<II", gz[at + bsize - 8:at + bsize])
        chunks.append((crc, isize))
        at += bsize
    assert at == len(gz)
    assert chunks[-1] == (0, 0) and gz[-28:] == BGZF_EOF
    want = [(zlib.crc32(text[i:i + 0xff00]) & 0xffffffff, len(text[i:i + 0xff00])) for i in range(0, len(text), 0xff00)]
    assert chunks[:-1] == want


@pytest.mark.parametrize("name", sorted(_texts()))
def test_bgzf_compress_round_trip(name):
    text = _texts()[name] * (3 if name in ("text", "acgt") else 1)
    gz = bgzf_compress_host(text)
    check_bgzf_layout(gz, text)
    assert gunzip_members(gz) == text                                 # zlib reads it
    rc, out = bgzf_inflate_host(gz, len(text))                        # and so does the decoder above
    assert rc == 0 and out == text
    if name == "text":
        assert len(gz) < len(text) // 2                               # VCF text: the LZ77 parse finds the repeats
    if name == "random":
        assert len(gz) <= len(text) + 5 + 26 + 28                     # incompressible: one stored member + EOF


@settings(max_examples=120, deadline=None)
@given(st.one_of(st.binary(max_size=3000), st.lists(st.sampled_from([b"A", b"C", b"chr1\t", b"0|1", b"\n", b"PASS", b"\xff\xfe"]), max_size=3000).map(b"".join)))
def test_bgzf_compress_property(data):
    gz = bgzf_compress_host(data)
    assert gunzip_members(gz) == data
    check_bgzf_layout(gz, data)


def test_bgzf_compress_long_matches_and_lengths():
    """every match length 3..258 and distances across the code boundaries: runs and periodic text"""
    rnd = random.Random(11)
    parts = []
    for length in list(range(3, 40)) + [57, 58, 59, 113, 114, 115, 130, 131, 226, 227, 257, 258, 259, 300]:
        seed = bytes(rnd.getrandbits(8) for _ in range(6))
        parts.append(seed + bytes(rnd.choice(b"ACGT") for _ in range(rnd.choice([1, 2, 5, 30, 200, 1500]))) + (seed * 60)[:length])
    text = b"".join(parts)
    assert gunzip_members(bgzf_compress_host(text)) == text
    far = bytes(rnd.getrandbits(8) for _ in range(300)) + bytes(rnd.choice(b"AC") for _ in range(32768 - 300)) + b"tail"
    far = far + far[:300]                                             # a repeat exactly 32768 back, and one just beyond the window
    assert gunzip_members(bgzf_compress_host(far)) == far


def _huff_lengths(freq, limit):
    f = np.asarray(freq, dtype=np.uint32)
    out = np.zeros(f.size, dtype=np.uint8)
    inf_lib().def_huff_lengths(f.ctypes.data_as(C.POINTER(C.c_uint32)), f.size, limit, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out.tolist()


def _optimal_cost(freq):
    import heapq
    h = [f for f in freq if f]
    if len(h) < 2:
        return sum(h)
    heapq.heapify(h)
    cost = 0
    while len(h) > 1:
        a, b = heapq.heappop(h), heapq.heappop(h)
        cost += a + b
        heapq.heappush(h, a + b)
    return cost


def test_huffman_lengths_are_prefix_codes_within_the_limit():
    rnd = random.Random(2)
    fib = [1, 1]
    while len(fib) < 40:
        fib.append(fib[-1] + fib[-2])
    cases = [[rnd.randint(0, 1000) for _ in range(286)], [rnd.choice([0, 0, 0, 5]) for _ in range(286)], fib[:30], fib[:19], fib[:25] + [0] * 200,
             [1] * 286, [0] * 29 + [7], [3, 0, 9], [2 ** k for k in range(19)], [0] * 30]
    for freq in cases:
        n = len(freq)
        limit = 7 if n == 19 else 15
        lens = _huff_lengths(freq, limit)
        used = [l for f, l in zip(freq, lens) if f]
        assert all(l == 0 for f, l in zip(freq, lens) if not f)
        assert all(1 <= l <= limit for l in used)
        if len(used) > 1:
            assert sum(2.0 ** -l for l in used) == 1.0                 # a complete prefix code (what zlib's inflate insists on)
        cost = sum(f * l for f, l in zip(freq, lens))
        opt = _optimal_cost(freq)
        if len(used) > 1 and max(_huff_lengths(freq, 64)) <= limit:
            assert cost == opt                                         # no limit in play: optimal
        elif len(used) > 1:
            assert opt <= cost <= 1.1 * opt + len(used)                # the halving heuristic costs little


def test_bgzf_compress_skewed_symbols_hit_the_length_limit():
    """byte k occurs F(k) times and no three-byte window repeats often enough to hide that: codes deeper than 15 bits are limited"""
    rnd = random.Random(9)
    fib = [1, 1]
    while len(fib) < 22:
        fib.append(fib[-1] + fib[-2])
    data = bytearray()
    for k, f in enumerate(fib):
        data += bytes([k + 65]) * f
    data = bytes(data)
    perm = list(data)
    rnd.shuffle(perm)
    for text in (data, bytes(perm)):
        gz = bgzf_compress_host(text)
        assert gunzip_members(gz) == text
        check_bgzf_layout(gz, text)
