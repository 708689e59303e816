"""ctypes binding of libsnpgpu.so (include/snpgpu.h).

This is the whole Python <-> CUDA boundary: plain pointers and sizes, no torch types.  There is no CPU
fallback -- if the shared library is missing or no CUDA device is visible the calls raise.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SNPGPU_LIB") or os.path.join(_HERE, "libsnpgpu.so")   # (override: tuning builds)

OK, E_VALUE, E_INDEX, E_UNPACK, E_DOMAIN, E_LONECR = 0, 1, 2, 3, 4, 5
E_CUDA, E_ARG, E_NOMEM, E_LENGTH = 16, 17, 18, 19
MODE_SITES, MODE_ALL = 0, 1
FAIL_RAWDPTH, FAIL_VARFREQ, FAIL_DEPTH, FAIL_STRDPTH, FAIL_STRBIAS, FAIL_REGION = 1, 2, 4, 8, 16, 32

EXPORTS = [
    "snpgpu_abi_version", "snpgpu_create", "snpgpu_destroy", "snpgpu_last_error", "snpgpu_set_stream",
    "snpgpu_sync", "snpgpu_host_alloc", "snpgpu_host_free", "snpgpu_launch_count", "snpgpu_enable_timing",
    "snpgpu_kernel_time", "snpgpu_sites_create", "snpgpu_sites_create_from_keys_dev",
    "snpgpu_sites_destroy", "snpgpu_sites_n_snp", "snpgpu_pileup_consensus", "snpgpu_pileup_consensus_dev",
    "snpgpu_pileup_consensus_begin", "snpgpu_pileup_consensus_end", "snpgpu_pileup_consensus_batch_dev",
    "snpgpu_normalize_newlines_dev", "snpgpu_pileup_vcf_records", "snpgpu_pileup_want_vcf_records", "snpgpu_reference_bases",
    "snpgpu_merge_sites", "snpgpu_merge_sites_dev", "snpgpu_pairwise_distance", "snpgpu_pairwise_distance_dev",
    "snpgpu_pairwise_distance_tiles_dev",
    "snpgpu_synth_pileup_dev", "snpgpu_synth_sample_sites", "snpgpu_pileup_depth_sum", "snpgpu_pileup_depth_sum_dev",
    "snpgpu_filter_regions", "snpgpu_pileup_vcf_text",
]


class Params(ctypes.Structure):
    """snpgpu_params: ConsensusCaller thresholds (pileup.py:433-471) + Reader's min_base_quality."""
    _fields_ = [("min_base_qual", ctypes.c_int32), ("min_cons_depth", ctypes.c_int32),
                ("min_cons_strand_depth", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("min_cons_freq", ctypes.c_double), ("min_cons_strand_bias", ctypes.c_double)]


class PileupStats(ctypes.Structure):
    _fields_ = [("n_lines", ctypes.c_uint64), ("n_parsed", ctypes.c_uint64), ("n_general", ctypes.c_uint64),
                ("error_offset", ctypes.c_uint64), ("error_code", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("n_called", ctypes.c_uint64)]


class PileupSample(ctypes.Structure):
    """snpgpu_pileup_sample: one sample of snpgpu_pileup_consensus_batch_dev (device pointers)."""
    _fields_ = [("text_dev", ctypes.c_void_p), ("nbytes", ctypes.c_size_t), ("row_out_dev", ctypes.c_void_p),
                ("line_out_dev", ctypes.c_void_p), ("line_out_cap", ctypes.c_size_t), ("stats_dev", ctypes.c_void_p)]


VCF_HAS_DEPTH, VCF_FIRST_IS_REF = 1, 2
VCF_FILTER_MASKS, VCF_FILTER_TEXT = 64, 64

# snpgpu_vcf_record / snpgpu_vcf_alt (include/snpgpu.h) as numpy record layouts
VCF_RECORD_DTYPE = np.dtype([("offset", "<u8"), ("pos", "<i8"), ("raw_depth", "<i8"), ("alt_index", "<u8"),
                             ("chrom_off", "<u4"), ("chrom_len", "<u4"), ("contig", "<i4"), ("rd", "<u4"),
                             ("rdf", "<u4"), ("rdr", "<u4"), ("n_alt", "<u4"), ("ref", "u1"), ("cons", "u1"),
                             ("fail", "u1"), ("flags", "u1")])
VCF_ALT_DTYPE = np.dtype([("ad", "<u4"), ("adf", "<u4"), ("adr", "<u4"), ("base", "u1"), ("pad", "u1", (3,))])
assert VCF_RECORD_DTYPE.itemsize == 64 and VCF_ALT_DTYPE.itemsize == 16


class SynthSpec(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("sample", ctypes.c_uint32), ("genome_len", ctypes.c_uint32),
                ("mean_depth", ctypes.c_uint32), ("n_pool_sites", ctypes.c_uint32),
                ("site_carry_prob", ctypes.c_float), ("indel_line_rate", ctypes.c_float)]


def make_params(min_base_qual=0, min_cons_freq=0.6, min_cons_depth=1, min_cons_strand_depth=0,
                min_cons_strand_bias=0.0):
    return Params(int(min_base_qual), int(min_cons_depth), int(min_cons_strand_depth), 0, float(min_cons_freq),
                  float(min_cons_strand_bias))


class SnpGpuError(RuntimeError):
    """A libsnpgpu call failed.  code is an SNPGPU_E_* value; codes 1-4 mean "the reference raises here"."""

    def __init__(self, code, message, offset=None):
        super().__init__("libsnpgpu error %d: %s" % (code, message))
        self.code = code
        self.offset = offset


_lib = None


def load():
    """Load libsnpgpu.so and declare every prototype of include/snpgpu.h.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SnpGpuError(E_CUDA, "%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "(there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, sz, i32, u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int32, ctypes.c_uint64
    P = ctypes.POINTER
    L.snpgpu_abi_version.restype = ctypes.c_int
    L.snpgpu_abi_version.argtypes = []
    L.snpgpu_create.restype = ctypes.c_int
    L.snpgpu_create.argtypes = [ctypes.c_int, P(vp)]
    L.snpgpu_destroy.restype = None
    L.snpgpu_destroy.argtypes = [vp]
    L.snpgpu_last_error.restype = ctypes.c_char_p
    L.snpgpu_last_error.argtypes = [vp]
    L.snpgpu_set_stream.restype = ctypes.c_int
    L.snpgpu_set_stream.argtypes = [vp, vp]
    L.snpgpu_sync.restype = ctypes.c_int
    L.snpgpu_sync.argtypes = [vp]
    L.snpgpu_host_alloc.restype = ctypes.c_int
    L.snpgpu_host_alloc.argtypes = [vp, sz, P(vp)]
    L.snpgpu_host_free.restype = ctypes.c_int
    L.snpgpu_host_free.argtypes = [vp, vp]
    L.snpgpu_launch_count.restype = u64
    L.snpgpu_launch_count.argtypes = [vp]
    L.snpgpu_enable_timing.restype = ctypes.c_int
    L.snpgpu_enable_timing.argtypes = [vp, ctypes.c_int]
    L.snpgpu_kernel_time.restype = ctypes.c_int
    L.snpgpu_kernel_time.argtypes = [vp, ctypes.c_int, P(ctypes.c_double), P(u64)]
    L.snpgpu_sites_create.restype = ctypes.c_int
    L.snpgpu_sites_create.argtypes = [vp, ctypes.c_char_p, vp, i32, vp, vp, sz, vp, vp, sz, P(vp)]
    L.snpgpu_sites_create_from_keys_dev.restype = ctypes.c_int
    L.snpgpu_sites_create_from_keys_dev.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_int32, vp, vp, sz, P(vp)]
    L.snpgpu_sites_destroy.restype = None
    L.snpgpu_sites_destroy.argtypes = [vp]
    L.snpgpu_sites_n_snp.restype = sz
    L.snpgpu_sites_n_snp.argtypes = [vp]
    L.snpgpu_pileup_consensus.restype = ctypes.c_int
    L.snpgpu_pileup_consensus.argtypes = [vp, vp, sz, vp, P(Params), ctypes.c_int, vp, vp, sz, P(PileupStats)]
    L.snpgpu_pileup_consensus_begin.restype = ctypes.c_int
    L.snpgpu_pileup_consensus_begin.argtypes = [vp, vp, sz, vp, P(Params), ctypes.c_int, vp, vp, sz, vp, P(ctypes.c_int)]
    L.snpgpu_pileup_consensus_end.restype = ctypes.c_int
    L.snpgpu_pileup_consensus_end.argtypes = [vp, ctypes.c_int]
    L.snpgpu_pileup_consensus_dev.restype = ctypes.c_int
    L.snpgpu_pileup_consensus_dev.argtypes = [vp, vp, sz, vp, P(Params), ctypes.c_int, vp, vp, sz, vp]
    L.snpgpu_pileup_consensus_batch_dev.restype = ctypes.c_int
    L.snpgpu_pileup_consensus_batch_dev.argtypes = [vp, P(PileupSample), sz, vp, P(Params), ctypes.c_int]
    L.snpgpu_normalize_newlines_dev.restype = ctypes.c_int
    L.snpgpu_normalize_newlines_dev.argtypes = [vp, vp, sz]
    L.snpgpu_reference_bases.restype = ctypes.c_int
    L.snpgpu_reference_bases.argtypes = [vp, vp, sz, vp, sz, vp, P(sz)]
    L.snpgpu_pileup_want_vcf_records.restype = ctypes.c_int
    L.snpgpu_pileup_want_vcf_records.argtypes = [vp, ctypes.c_int]
    L.snpgpu_pileup_vcf_records.restype = ctypes.c_int
    L.snpgpu_pileup_vcf_records.argtypes = [vp, vp, P(Params), ctypes.c_int, vp, sz, P(sz), vp, sz, P(sz)]
    L.snpgpu_merge_sites.restype = ctypes.c_int
    L.snpgpu_merge_sites.argtypes = [vp, vp, vp, sz, vp, vp, vp, P(sz)]
    L.snpgpu_merge_sites_dev.restype = ctypes.c_int
    L.snpgpu_merge_sites_dev.argtypes = [vp, vp, vp, sz, vp, vp, vp, P(sz)]
    L.snpgpu_pairwise_distance.restype = ctypes.c_int
    L.snpgpu_pairwise_distance.argtypes = [vp, vp, sz, sz, sz, vp]
    L.snpgpu_pairwise_distance_dev.restype = ctypes.c_int
    L.snpgpu_pairwise_distance_dev.argtypes = [vp, vp, sz, sz, sz, sz, sz, vp]
    L.snpgpu_pairwise_distance_tiles_dev.restype = ctypes.c_int
    L.snpgpu_pairwise_distance_tiles_dev.argtypes = [vp, vp, sz, sz, sz, vp, sz, vp]
    L.snpgpu_pileup_vcf_text.restype = ctypes.c_int
    L.snpgpu_pileup_vcf_text.argtypes = [vp, vp, vp, i32, ctypes.c_char_p, i32, i32, vp, sz, ctypes.POINTER(sz), ctypes.POINTER(sz)]
    L.snpgpu_filter_regions.restype = ctypes.c_int
    L.snpgpu_filter_regions.argtypes = [vp, vp, vp, sz, vp, vp, i32, vp, vp, sz, vp]
    L.snpgpu_pileup_depth_sum.restype = ctypes.c_int
    L.snpgpu_pileup_depth_sum.argtypes = [vp, vp, sz, P(ctypes.c_int64), P(u64), P(u64)]
    L.snpgpu_pileup_depth_sum_dev.restype = ctypes.c_int
    L.snpgpu_pileup_depth_sum_dev.argtypes = [vp, vp, sz, P(ctypes.c_int64), P(u64), P(u64)]
    L.snpgpu_synth_pileup_dev.restype = ctypes.c_int
    L.snpgpu_synth_pileup_dev.argtypes = [vp, P(SynthSpec), ctypes.c_char_p, vp, sz, P(sz)]
    L.snpgpu_synth_sample_sites.restype = ctypes.c_int
    L.snpgpu_synth_sample_sites.argtypes = [vp, P(SynthSpec), vp, sz, P(sz)]
    _lib = L
    return L


def _np_ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None and a.size else None


class Sites(object):
    """Device-resident site table: the snplist entries (in file order) + the exclude VCF's positions."""

    def __init__(self, ctx, snp_list, excluded=()):
        self.ctx = ctx
        excluded = list(excluded)
        names = sorted({c for c, _ in snp_list} | {c for c, _ in excluded})
        self.contigs = names
        idx = {c: i for i, c in enumerate(names)}
        enc = [n.encode("utf-8") for n in names]
        blob = b"".join(enc)
        off = np.zeros(len(names) + 1, dtype=np.int32)
        if names:
            off[1:] = np.cumsum([len(n) for n in enc])
        sc = np.array([idx[c] for c, _ in snp_list], dtype=np.int32)
        sp = np.array([p for _, p in snp_list], dtype=np.int64)
        ec = np.array([idx[c] for c, _ in excluded], dtype=np.int32)
        ep = np.array([p for _, p in excluded], dtype=np.int64)
        self.n_snp = len(snp_list)
        h = ctypes.c_void_p()
        rc = ctx.lib.snpgpu_sites_create(ctx.handle, blob, _np_ptr(off), len(names), _np_ptr(sc), _np_ptr(sp),
                                         len(snp_list), _np_ptr(ec), _np_ptr(ep), len(excluded), ctypes.byref(h))
        ctx._check(rc)
        self.handle = h

    @classmethod
    def from_keys_dev(cls, ctx, contigs, contig_len, keys_ptr, n_keys):
        """Device fast path: the table straight from merge_sites_dev's sorted unique keys (device pointer), no host
        round trip.  contigs: names in chromosome-string order (index = chrom_rank); contig_len: their lengths."""
        self = cls.__new__(cls)
        self.ctx = ctx
        self.contigs = list(contigs)
        enc = [n.encode("utf-8") for n in self.contigs]
        blob = b"".join(enc)
        off = np.zeros(len(enc) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(n) for n in enc])
        lens = np.ascontiguousarray(contig_len, dtype=np.int64)
        self.n_snp = int(n_keys)
        h = ctypes.c_void_p()
        rc = ctx.lib.snpgpu_sites_create_from_keys_dev(ctx.handle, blob, _np_ptr(off), len(enc), _np_ptr(lens),
                                                       ctypes.c_void_p(int(keys_ptr)), int(n_keys), ctypes.byref(h))
        ctx._check(rc)
        self.handle = h
        return self

    @classmethod
    def from_arrays(cls, ctx, contigs, snp_contig, snp_pos, exc_contig=None, exc_pos=None):
        """Array fast path: contigs = list of names; snp_contig int32 indices into it, snp_pos int64."""
        self = cls.__new__(cls)
        self.ctx = ctx
        self.contigs = list(contigs)
        enc = [n.encode("utf-8") for n in self.contigs]
        blob = b"".join(enc)
        off = np.zeros(len(enc) + 1, dtype=np.int32)
        if enc:
            off[1:] = np.cumsum([len(n) for n in enc])
        sc = np.ascontiguousarray(snp_contig, dtype=np.int32)
        sp = np.ascontiguousarray(snp_pos, dtype=np.int64)
        ec = np.ascontiguousarray(exc_contig if exc_contig is not None else [], dtype=np.int32)
        ep = np.ascontiguousarray(exc_pos if exc_pos is not None else [], dtype=np.int64)
        self.n_snp = int(sp.size)
        h = ctypes.c_void_p()
        rc = ctx.lib.snpgpu_sites_create(ctx.handle, blob, _np_ptr(off), len(enc), _np_ptr(sc), _np_ptr(sp), sp.size,
                                         _np_ptr(ec), _np_ptr(ep), ep.size, ctypes.byref(h))
        ctx._check(rc)
        self.handle = h
        return self

    def close(self):
        if self.handle:
            self.ctx.lib.snpgpu_sites_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class Context(object):
    """One libsnpgpu context: one GPU, one stream.  Not thread-safe (snpgpu.h)."""

    def __init__(self, device=0):
        self.lib = load()
        if self.lib.snpgpu_abi_version() != 2:
            raise SnpGpuError(E_ARG, "libsnpgpu ABI mismatch")
        h = ctypes.c_void_p()
        rc = self.lib.snpgpu_create(int(device), ctypes.byref(h))
        if rc:
            raise SnpGpuError(rc, "snpgpu_create(device=%d) failed: no usable CUDA device "
                                  "(libsnpgpu has no CPU fallback)" % device)
        self.handle = h
        self.device = device

    # -- plumbing ---------------------------------------------------------------------------------
    def _check(self, rc, offset=None):
        if rc:
            msg = self.lib.snpgpu_last_error(self.handle)
            raise SnpGpuError(rc, msg.decode("utf-8", "replace") if msg else "", offset)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.snpgpu_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def set_stream(self, cuda_stream):
        self._check(self.lib.snpgpu_set_stream(self.handle, ctypes.c_void_p(int(cuda_stream) if cuda_stream else 0)))

    def sync(self):
        self._check(self.lib.snpgpu_sync(self.handle))

    @property
    def launch_count(self):
        return int(self.lib.snpgpu_launch_count(self.handle))

    def enable_timing(self, on=True):
        self._check(self.lib.snpgpu_enable_timing(self.handle, 1 if on else 0))

    def kernel_time(self, kernel):
        """(summed device ms, launches) of kernel 0 = pileup, 1 = distance since the last call; synchronises."""
        ms, n = ctypes.c_double(0), ctypes.c_uint64(0)
        self._check(self.lib.snpgpu_kernel_time(self.handle, int(kernel), ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, int(n.value)

    def pinned_array(self, nbytes):
        """(uint8 ndarray over page-locked host memory, owner); call owner.free() when done with the array."""
        p = ctypes.c_void_p()
        self._check(self.lib.snpgpu_host_alloc(self.handle, int(nbytes), ctypes.byref(p)))
        buf = (ctypes.c_uint8 * max(int(nbytes), 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.uint8, count=int(nbytes))
        return arr, _PinnedOwner(self, p)

    def sites(self, snp_list, excluded=()):
        return Sites(self, snp_list, excluded)

    # -- K1 ---------------------------------------------------------------------------------------
    def pileup_consensus(self, text, sites, params, mode=MODE_SITES, want_lines=False):
        """Host-buffer entry point.  text: bytes-like / uint8 ndarray with the pileup file's contents.
        Returns (row bytes, stats[, per-line uint16 array in file order])."""
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else text
        if buf.dtype != np.uint8 or not buf.flags["C_CONTIGUOUS"]:
            raise SnpGpuError(E_ARG, "text must be a contiguous uint8 array")
        row = np.zeros(max(sites.n_snp, 1), dtype=np.uint8)
        stats = PileupStats()
        lines = None
        cap = 0
        if want_lines and mode == MODE_ALL:
            cap = int(buf.size // 4 + 2)           # a line has at least 4 bytes ("c 1\n")
            lines = np.zeros(cap, dtype=np.uint16)
        rc = self.lib.snpgpu_pileup_consensus(self.handle, _np_ptr(buf), buf.size, sites.handle,
                                              ctypes.byref(params), mode, _np_ptr(row), _np_ptr(lines), cap,
                                              ctypes.byref(stats))
        self._check(rc, stats.error_offset)
        out_row = row[:sites.n_snp].tobytes()
        if lines is not None:
            return out_row, stats, lines[:stats.n_lines]
        return out_row, stats

    def reference_bases(self, seq, pos):
        """upper(seq[pos - 1]) for every 1-based position, with Python's indexing (snpgpu_reference_bases).
        seq: uint8 array, pos: int64 array.  Raises IndexError where the reference's indexing does."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        pos = np.ascontiguousarray(pos, dtype=np.int64)
        out = np.zeros(pos.size, dtype=np.uint8)
        bad = ctypes.c_size_t(0)
        rc = self.lib.snpgpu_reference_bases(self.handle, _np_ptr(seq), seq.size, _np_ptr(pos), pos.size, _np_ptr(out),
                                             ctypes.byref(bad))
        if rc == E_INDEX:
            raise IndexError("index out of range")             # what Bio.Seq / str indexing raises (utils.py:1108)
        self._check(rc)
        return out

    def want_vcf_records(self, on=True):
        """pileup_vcf_records() will follow the next pileup_consensus() calls: list the parsed lines on the way."""
        self._check(self.lib.snpgpu_pileup_want_vcf_records(self.handle, 1 if on else 0))

    def pileup_vcf_records(self, sites, params, mode=MODE_SITES):
        """K5: the tallies behind the consensus VCF, one record per pileup line that the preceding pileup_consensus()
        call parsed (same sites / params / mode), in file order.  Returns (records, alts) as numpy record arrays
        (VCF_RECORD_DTYPE, VCF_ALT_DTYPE); a record's ALT alleles are alts[alt_index : alt_index + n_alt]."""
        rec_cap, alt_cap = 4 * sites.n_snp + 1024, 8 * sites.n_snp + 2048
        for _ in range(3):
            rec = np.zeros(rec_cap, dtype=VCF_RECORD_DTYPE)
            alt = np.zeros(alt_cap, dtype=VCF_ALT_DTYPE)
            n_rec, n_alt = ctypes.c_size_t(0), ctypes.c_size_t(0)
            rc = self.lib.snpgpu_pileup_vcf_records(self.handle, sites.handle, ctypes.byref(params), mode, _np_ptr(rec),
                                                    rec_cap, ctypes.byref(n_rec), _np_ptr(alt), alt_cap,
                                                    ctypes.byref(n_alt))
            if rc == E_NOMEM and (n_rec.value > rec_cap or n_alt.value > alt_cap):
                rec_cap, alt_cap = max(rec_cap, n_rec.value), max(alt_cap, n_alt.value)
                continue
            self._check(rc)
            return rec[:n_rec.value], alt[:n_alt.value]
        raise SnpGpuError(E_NOMEM, "pileup_vcf_records: capacities kept growing")

    def pileup_vcf_text(self, sites, params, mode, filter_texts, failed_snp_gt=".", preserve_ref_case=False):
        """K5 with the formatting on the device: the data lines of the consensus VCF of the preceding pileup_consensus()
        call as a uint8 array, and their number.  filter_texts: the FILTER column's text for every fail mask 0 .. 63."""
        assert len(filter_texts) == VCF_FILTER_MASKS
        table = bytearray(VCF_FILTER_MASKS * VCF_FILTER_TEXT)
        for m, t in enumerate(filter_texts):
            b = t.encode("ascii")
            if len(b) >= VCF_FILTER_TEXT:
                raise SnpGpuError(E_ARG, "pileup_vcf_text: filter text longer than %d bytes" % (VCF_FILTER_TEXT - 1))
            table[m * VCF_FILTER_TEXT:m * VCF_FILTER_TEXT + len(b)] = b
        cap = 1 << 20
        gt = ord(failed_snp_gt) if failed_snp_gt in (".", "0") else ord("1")
        for _ in range(3):
            out = np.empty(cap, dtype=np.uint8)
            n_text, n_rec = ctypes.c_size_t(0), ctypes.c_size_t(0)
            rc = self.lib.snpgpu_pileup_vcf_text(self.handle, sites.handle, ctypes.byref(params), mode, bytes(table), gt,
                                                 1 if preserve_ref_case else 0, _np_ptr(out), cap, ctypes.byref(n_text),
                                                 ctypes.byref(n_rec))
            if rc == E_NOMEM and n_text.value > cap:
                cap = n_text.value
                continue
            self._check(rc)
            return out[:n_text.value], n_rec.value
        raise SnpGpuError(E_NOMEM, "pileup_vcf_text: capacity kept growing")

    def pileup_consensus_begin(self, text, sites, params, mode, row_out, line_out=None, stats=None):
        """First half of the pipelined host-buffer call (snpgpu_pileup_consensus_begin): text / row_out / line_out are
        numpy arrays (page-locked ones from pinned_array() for real overlap), stats a PileupStats; all must stay alive
        until pileup_consensus_end(slot).  Returns the slot."""
        slot = ctypes.c_int(-1)
        rc = self.lib.snpgpu_pileup_consensus_begin(
            self.handle, _np_ptr(text), text.size, sites.handle, ctypes.byref(params), mode, _np_ptr(row_out),
            _np_ptr(line_out) if line_out is not None else None, line_out.size if line_out is not None else 0,
            ctypes.byref(stats) if stats is not None else None, ctypes.byref(slot))
        self._check(rc)
        return slot.value

    def pileup_consensus_end(self, slot):
        self._check(self.lib.snpgpu_pileup_consensus_end(self.handle, slot))

    def pileup_consensus_dev(self, text_ptr, nbytes, sites, params, mode, row_ptr, line_ptr=0, line_cap=0,
                             stats_ptr=0):
        """Device-pointer entry point (nothing is synchronised)."""
        rc = self.lib.snpgpu_pileup_consensus_dev(self.handle, ctypes.c_void_p(text_ptr), int(nbytes), sites.handle,
                                                  ctypes.byref(params), mode, ctypes.c_void_p(row_ptr or 0),
                                                  ctypes.c_void_p(line_ptr or 0), int(line_cap),
                                                  ctypes.c_void_p(stats_ptr or 0))
        self._check(rc)

    def pileup_consensus_batch_dev(self, samples, sites, params, mode):
        """Batch form of pileup_consensus_dev: samples = sequence of (text_ptr, nbytes, row_ptr, line_ptr, line_cap,
        stats_ptr) tuples of device pointers (0 = none), or a ready (PileupSample * n) array.  One launch sequence per up to 64
        samples; nothing is synchronised."""
        if isinstance(samples, ctypes.Array):
            arr = samples
        else:
            arr = (PileupSample * len(samples))()
            for k, (tp, nb, rp, lp, lc, sp) in enumerate(samples):
                arr[k] = PileupSample(tp or None, int(nb), rp or None, lp or None, int(lc), sp or None)
        self._check(self.lib.snpgpu_pileup_consensus_batch_dev(self.handle, arr, len(arr), sites.handle,
                                                               ctypes.byref(params), mode))

    # -- K7 ---------------------------------------------------------------------------------------
    def filter_regions(self, snp_keys, seg_last, max_snps, window, edge_keys, edge_end):
        """removed[i] for SNP i (snpgpu_filter_regions): keys = group << 48 | contig rank << 32 | pos, segments sorted."""
        k = np.ascontiguousarray(snp_keys, dtype=np.uint64)
        last = np.ascontiguousarray(seg_last, dtype=np.uint32)
        mx = np.ascontiguousarray(max_snps, dtype=np.int32)
        win = np.ascontiguousarray(window, dtype=np.int32)
        ek = np.ascontiguousarray(edge_keys, dtype=np.uint64)
        ee = np.ascontiguousarray(edge_end, dtype=np.uint32)
        out = np.zeros(k.size, dtype=np.uint8)
        self._check(self.lib.snpgpu_filter_regions(self.handle, _np_ptr(k), _np_ptr(last), k.size, _np_ptr(mx), _np_ptr(win),
                                                   mx.size, _np_ptr(ek), _np_ptr(ee), ek.size, _np_ptr(out)))
        return out.astype(bool)

    # -- K6 ---------------------------------------------------------------------------------------
    def pileup_depth_sum(self, text):
        """(sum of the raw-depth column over all lines, lines that contributed): collect_metrics.py:322-329.
        text: bytes-like / uint8 ndarray with the pileup file's contents."""
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else text
        total, lines, off = ctypes.c_int64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
        rc = self.lib.snpgpu_pileup_depth_sum(self.handle, _np_ptr(buf), buf.size, ctypes.byref(total), ctypes.byref(lines),
                                              ctypes.byref(off))
        self._check(rc, off.value)
        return total.value, lines.value

    # -- K2 ---------------------------------------------------------------------------------------
    def merge_sites(self, keys, sample_of):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        sample_of = np.ascontiguousarray(sample_of, dtype=np.uint32)
        n = keys.size
        uniq = np.zeros(max(n, 1), dtype=np.uint64)
        cnt = np.zeros(max(n, 1), dtype=np.uint32)
        samples = np.zeros(max(n, 1), dtype=np.uint32)
        nu = ctypes.c_size_t(0)
        self._check(self.lib.snpgpu_merge_sites(self.handle, _np_ptr(keys), _np_ptr(sample_of), n, _np_ptr(uniq),
                                                _np_ptr(cnt), _np_ptr(samples), ctypes.byref(nu)))
        return uniq[:nu.value], cnt[:nu.value], samples[:n]

    def merge_sites_dev(self, keys_ptr, sample_ptr, n, uniq_ptr, cnt_ptr, samples_ptr):
        nu = ctypes.c_size_t(0)
        self._check(self.lib.snpgpu_merge_sites_dev(self.handle, ctypes.c_void_p(keys_ptr), ctypes.c_void_p(sample_ptr),
                                                    int(n), ctypes.c_void_p(uniq_ptr), ctypes.c_void_p(cnt_ptr),
                                                    ctypes.c_void_p(samples_ptr), ctypes.byref(nu)))
        return nu.value

    # -- K4 ---------------------------------------------------------------------------------------
    def pairwise_distance(self, matrix):
        """matrix: uint8 [n_rows, n_sites] -> int32 [n_rows, n_rows]."""
        m = np.ascontiguousarray(matrix, dtype=np.uint8)
        n, s = (m.shape + (0,))[:2] if m.ndim == 2 else (0, 0)
        d = np.zeros((n, n), dtype=np.int32)
        if n:
            self._check(self.lib.snpgpu_pairwise_distance(self.handle, _np_ptr(m), n, s, s, _np_ptr(d)))
        return d

    def pairwise_distance_dev(self, matrix_ptr, n_rows, n_sites, row_stride, row_begin, row_end, dist_ptr):
        self._check(self.lib.snpgpu_pairwise_distance_dev(self.handle, ctypes.c_void_p(matrix_ptr), int(n_rows),
                                                          int(n_sites), int(row_stride), int(row_begin),
                                                          int(row_end), ctypes.c_void_p(dist_ptr)))

    def pairwise_distance_tiles_dev(self, matrix_ptr, n_rows, n_sites, row_stride, tile_rows, dist_ptr):
        """the 64-row tile rows listed in tile_rows, each from its diagonal tile rightwards, 64 output rows per entry"""
        t = np.ascontiguousarray(tile_rows, dtype=np.uint32)
        self._check(self.lib.snpgpu_pairwise_distance_tiles_dev(self.handle, ctypes.c_void_p(matrix_ptr), int(n_rows),
                                                                int(n_sites), int(row_stride), _np_ptr(t), t.size,
                                                                ctypes.c_void_p(dist_ptr)))

    # -- synthetic input (bench / tests) ------------------------------------------------------------
    def synth_pileup_dev(self, spec, contig_name, text_ptr, cap):
        n = ctypes.c_size_t(0)
        self._check(self.lib.snpgpu_synth_pileup_dev(self.handle, ctypes.byref(spec), contig_name.encode(),
                                                     ctypes.c_void_p(text_ptr), int(cap), ctypes.byref(n)))
        return n.value

    def synth_sample_sites(self, spec):
        cap = int(spec.n_pool_sites * max(float(spec.site_carry_prob), 0.0) * 2 + 4096)
        while True:
            pos = np.zeros(cap, dtype=np.uint32)
            n = ctypes.c_size_t(0)
            self._check(self.lib.snpgpu_synth_sample_sites(self.handle, ctypes.byref(spec), _np_ptr(pos), cap,
                                                           ctypes.byref(n)))
            if n.value <= cap:
                return pos[:n.value].copy()
            cap = n.value


class _PinnedOwner(object):
    def __init__(self, ctx, ptr):
        self.ctx, self.ptr = ctx, ptr

    def free(self):
        if self.ptr is not None and self.ctx.handle:
            self.ctx.lib.snpgpu_host_free(self.ctx.handle, self.ptr)
        self.ptr = None
