"""VCF ingest on the device: `parse_vcf_text` hands inflated VCF record lines to avk_vcf_parse (parse_variant /
parse_genotype / get_variant_type of src/parsing/region_generation.rs:565-758, one record per thread) and returns the records
in the generator's call-set form.  BGZF files are inflated on the device too (`bgzf_inflate`, `parse_vcf_bgzf`: one thread per
BGZF member, avk_inflate.cuh); `read_vcf_lines` (Python's gzip) remains for plain gzip / uncompressed files and as the stand-in
for noodles' tabix query."""
import ctypes as C
import gzip

import numpy as np

from . import abi


def read_vcf_lines(path):
    """Whole (b)gzipped or plain VCF -> bytes of its lines (header included; the device parser skips '#' lines)."""
    with open(path, "rb") as f:
        magic = f.read(2)
    opener = gzip.open if magic == b"\x1f\x8b" else open
    with opener(path, "rb") as f:
        return f.read()


class VcfTable:
    """One input of an avk_callsets table plus its variant_contig array."""

    def __init__(self, cap_variants, cap_pool):
        n, p = max(int(cap_variants), 1), max(int(cap_pool), 1)
        self.contig = np.zeros(n, np.uint32); self.position = np.zeros(n, np.uint32)
        self.variant_type = np.zeros(n, np.uint8); self.zygosity = np.zeros(n, np.uint8)
        self.raw_allele_space = np.zeros(n, np.uint32); self.allele_off = np.zeros(n, np.uint32)
        self.a0_len = np.zeros(n, np.uint32); self.a1_len = np.zeros(n, np.uint32)
        self.allele_pool = np.zeros(p, np.uint8)
        self.n_variants = 0
        self.pool_len = 0

    def to_c(self):
        return abi.VcfOut(0, 0, self.contig.size, self.allele_pool.size, abi.ptr(self.contig), abi.ptr(self.position), abi.ptr(self.variant_type),
                          abi.ptr(self.zygosity), abi.ptr(self.raw_allele_space), abi.ptr(self.allele_off), abi.ptr(self.a0_len), abi.ptr(self.a1_len),
                          abi.ptr(self.allele_pool))

    def finish(self, c):
        self.n_variants, self.pool_len = int(c.n_variants), int(c.allele_pool_len)
        return self

    def records(self):
        """[(contig, position, allele0, allele1, zygosity, type, raw_allele_space)] -- the generator's record form with the contig first."""
        out = []
        for v in range(self.n_variants):
            o, l0, l1 = int(self.allele_off[v]), int(self.a0_len[v]), int(self.a1_len[v])
            out.append((int(self.contig[v]), int(self.position[v]), self.allele_pool[o:o + l0].tobytes(), self.allele_pool[o + l0:o + l0 + l1].tobytes(),
                        int(self.zygosity[v]), int(self.variant_type[v]), int(self.raw_allele_space[v])))
        return out


def parse_vcf_text(solver, text: bytes, contig_names, sample_index=0, enable_trimming=True) -> VcfTable:
    """avk_vcf_parse through a Solver's context.  Raises AvkError naming the first record the reference would fail on."""
    lib = solver._lib
    lib.avk_vcf_parse.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32, C.c_int, C.POINTER(abi.VcfOut),
                                  C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
    tab = VcfTable(2 * (text.count(b"\n") + 1), len(text))
    c = tab.to_c()
    names = (C.c_char_p * len(contig_names))(*[n.encode() for n in contig_names])
    line, code = C.c_uint64(0), C.c_int32(0)
    solver._check(lib.avk_vcf_parse(solver._ctx, text, len(text), names, len(contig_names), sample_index, 1 if enable_trimming else 0, C.byref(c),
                                    C.byref(line), C.byref(code)), "avk_vcf_parse")
    return tab.finish(c)


def bgzf_inflate(solver, gz: bytes, verify_crc=True) -> bytes:
    """avk_bgzf_inflate: a whole BGZF file -> its bytes, inflated on the device (raises AvkError for plain gzip / damaged input)."""
    lib = solver._lib
    lib.avk_bgzf_inflate.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
    n = C.c_uint64(0)
    solver._check(lib.avk_bgzf_inflate(solver._ctx, gz, len(gz), 1 if verify_crc else 0, None, 0, C.byref(n)), "avk_bgzf_inflate")
    buf = C.create_string_buffer(max(int(n.value), 1))
    solver._check(lib.avk_bgzf_inflate(solver._ctx, gz, len(gz), 1 if verify_crc else 0, buf, int(n.value), C.byref(n)), "avk_bgzf_inflate")
    return buf.raw[:int(n.value)]


def parse_vcf_bgzf(solver, gz: bytes, contig_names, sample_index=0, enable_trimming=True, verify_crc=True) -> VcfTable:
    """avk_vcf_parse_bgzf: BGZF-compressed VCF -> call-set table; inflate and parse both run on the device."""
    lib = solver._lib
    lib.avk_bgzf_inflate.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.avk_vcf_parse_bgzf.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32, C.c_int, C.POINTER(abi.VcfOut),
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
    n = C.c_uint64(0)
    solver._check(lib.avk_bgzf_inflate(solver._ctx, gz, len(gz), 0, None, 0, C.byref(n)), "avk_bgzf_inflate")     # inflated size: bounds the table
    size = int(n.value)
    tab = VcfTable(size // 16 + 2, size)          # first guess; the call reports the sizes it needs when this is too small
    c = tab.to_c()
    names = (C.c_char_p * len(contig_names))(*[x.encode() for x in contig_names])
    line, code = C.c_uint64(0), C.c_int32(0)
    rc = lib.avk_vcf_parse_bgzf(solver._ctx, gz, len(gz), 1 if verify_crc else 0, names, len(contig_names), sample_index, 1 if enable_trimming else 0,
                                C.byref(c), C.byref(line), C.byref(code))
    if rc == -4 and (int(c.n_variants) > tab.contig.size or int(c.allele_pool_len) > tab.allele_pool.size):   # AVK_ERR_OOM: the call reports what it needs
        tab = VcfTable(int(c.n_variants), int(c.allele_pool_len))
        c = tab.to_c()
        rc = lib.avk_vcf_parse_bgzf(solver._ctx, gz, len(gz), 1 if verify_crc else 0, names, len(contig_names), sample_index, 1 if enable_trimming else 0,
                                    C.byref(c), C.byref(line), C.byref(code))
    solver._check(rc, "avk_vcf_parse_bgzf")
    return tab.finish(c)
