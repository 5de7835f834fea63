#!/usr/bin/env python
"""Timing of the device BGZF inflate + VCF parse on a synthetic VCF of about N records (default 1 M), against zlib on one host core."""
import json, os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    import test_bgzf as B
    from aardvark_b200.ingest import bgzf_inflate, parse_vcf_bgzf, parse_vcf_text
    from aardvark_b200.lib import Solver
    rng = np.random.default_rng(5)
    pos = np.sort(rng.integers(1, 240_000_000, n))
    bases = [b"A", b"C", b"G", b"T"]
    ref = rng.integers(0, 4, n); alt = (ref + rng.integers(1, 4, n)) % 4
    gts = [b"0/1", b"1/1", b"0|1", b"1|0"]
    gt = rng.integers(0, 4, n)
    lines = [b"chr1\t%d\t.\t%s\t%s\t%d\tPASS\tDP=%d\tGT:GQ:DP\t%s:%d:%d\n" % (pos[i], bases[ref[i]], bases[alt[i]], 30 + i % 40, 20 + i % 30, gts[gt[i]], 20 + i % 70, 10 + i % 50)
             for i in range(n)]
    text = b"##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\n" + b"".join(lines)
    t0 = time.perf_counter(); gz = B.bgzf_compress(text, 6); t_comp = time.perf_counter() - t0
    t0 = time.perf_counter(); d = zlib.decompressobj(31); out = []
    buf = gz
    while buf:
        out.append(d.decompress(buf)); buf = d.unused_data
        if buf:
            d = zlib.decompressobj(31)
    t_zlib = time.perf_counter() - t0
    assert b"".join(out) == text
    s = Solver(0)
    for _ in range(2):
        got = bgzf_inflate(s, gz)
    assert got == text
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        bgzf_inflate(s, gz)
    t_inf = (time.perf_counter() - t0) / reps
    names = ["chr1"]
    tab = parse_vcf_bgzf(s, gz, names)
    t0 = time.perf_counter()
    for _ in range(reps):
        tab = parse_vcf_bgzf(s, gz, names)
    t_parse = (time.perf_counter() - t0) / reps
    assert tab.n_variants == n
    from aardvark_b200.writers import bgzf_compress
    ours = bgzf_compress(s, text)
    assert B.gunzip_members(ours) == text
    t0 = time.perf_counter()
    for _ in range(reps):
        bgzf_compress(s, text)
    t_def = (time.perf_counter() - t0) / reps
    print(json.dumps({"records": n, "text_bytes": len(text), "bgzf_bytes": len(gz), "members": (len(text) + 0xfeff) // 0xff00,
                      "zlib_one_core_ms": t_zlib * 1e3, "device_inflate_to_host_ms": t_inf * 1e3, "device_inflate_plus_parse_ms": t_parse * 1e3,
                      "inflate_MBps_out": len(text) / t_inf / 1e6,
                      "device_compress_ms": t_def * 1e3, "device_compress_bytes": len(ours), "zlib6_compress_one_core_ms": t_comp * 1e3, "note": "wall clock through the Python binding, H2D of the file and D2H of the text / table included"}))


if __name__ == "__main__":
    main()
