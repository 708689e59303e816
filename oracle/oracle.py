"""CPU oracle for the pileup -> consensus -> SNP-matrix -> distance hot path.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py; the product package never imports it.

The arithmetic lives in ``snp_oracle.c`` (compiled here with gcc into ``oracle/_build/liboracle.so``);
this module adds the file-level behaviour of the reference's four subcommands as small pure-Python
functions, each citing the reference lines it restates (paths relative to the upstream repository).

Third-party pieces the reference leans on and that are NOT in its tree (named + restated here):
  * PyVCF3 ~=1.0.3 (setup.py:25) -- ``vcf.Reader`` as used at utils.py:1127: only CHROM and POS are consumed.
    Restated in :func:`vcf_sites`; pinned by the bundled snplist*.txt golden files (lambda/agona/listeria).
  * Biopython (setup.py:28) -- ``SeqIO.write(..., "fasta")`` as called at call_consensus.py:189-192.
    Restated in :func:`fasta_text`; pinned by every bundled consensus*.fasta / snpma*.fasta.
Parity status: PINNED (tests/test_oracle_golden.py, tests/test_oracle_vs_reference.py).
"""
from __future__ import annotations

import ctypes
import itertools
import os
import re
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "snp_oracle.c")
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "liboracle.so")

FAIL_RAWDPTH, FAIL_VARFREQ, FAIL_DEPTH, FAIL_STRDPTH, FAIL_STRBIAS, FAIL_REGION = 1, 2, 4, 8, 16, 32
E_OK, E_VALUE, E_INDEX, E_UNPACK, E_DOMAIN = 0, 1, 2, 3, 4


class OracleError(Exception):
    """The reference would raise on this input (status = ORACLE_E_*, line = 0-based line index)."""

    def __init__(self, status, line):
        super().__init__("oracle: reference raises (status %d) at line %d" % (status, line))
        self.status = status
        self.line = line


class Params(ctypes.Structure):
    _fields_ = [("min_base_qual", ctypes.c_int), ("min_cons_freq", ctypes.c_double),
                ("min_cons_depth", ctypes.c_int), ("min_cons_strand_depth", ctypes.c_int),
                ("min_cons_strand_bias", ctypes.c_double)]


def make_params(min_base_qual=0, min_cons_freq=0.6, min_cons_depth=1, min_cons_strand_depth=0,
                min_cons_strand_bias=0.0):
    return Params(int(min_base_qual), float(min_cons_freq), int(min_cons_depth), int(min_cons_strand_depth),
                  float(min_cons_strand_bias))


def build(force=False):
    """Compile snp_oracle.c (gcc -O3).  Returns the path of the shared object."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(_BUILD, exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-std=c11", "-fPIC", "-shared", "-o", _SO, _SRC])
    return _SO


# ---- the bench's synthetic input written on the host (oracle/synth_host.cpp: the generator's own line function,
#      snp_pipeline_b200/csrc/synth_line.cuh, compiled for the host -- bench.py --impl reference must not load the GPU library)
_SYNTH_SRC = os.path.join(_HERE, "synth_host.cpp")
_SYNTH_HDRS = [os.path.join(_HERE, "..", "snp_pipeline_b200", "csrc", n) for n in ("synth_line.cuh", "hd.cuh")]
_SYNTH_SO = os.path.join(_BUILD, "libsynthhost.so")
_synth = None


class SynthSpec(ctypes.Structure):                           # = snpgpu_synth_spec (include/snpgpu.h)
    _fields_ = [("seed", ctypes.c_uint64), ("sample", ctypes.c_uint32), ("genome_len", ctypes.c_uint32),
                ("mean_depth", ctypes.c_uint32), ("n_pool_sites", ctypes.c_uint32),
                ("site_carry_prob", ctypes.c_float), ("indel_line_rate", ctypes.c_float)]


def build_synth(force=False):
    """Compile synth_host.cpp (g++ -O2).  Returns the path of the shared object."""
    srcs = [_SYNTH_SRC] + [h for h in _SYNTH_HDRS if os.path.exists(h)]
    if force or not os.path.exists(_SYNTH_SO) or os.path.getmtime(_SYNTH_SO) < max(os.path.getmtime(s) for s in srcs):
        os.makedirs(_BUILD, exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", _SYNTH_SO, _SYNTH_SRC])
    return _SYNTH_SO


def _synth_lib():
    global _synth
    if _synth is None:
        L = ctypes.CDLL(build_synth())
        L.synth_host_pileup.restype = ctypes.c_ulonglong
        L.synth_host_pileup.argtypes = [ctypes.POINTER(SynthSpec), ctypes.c_char_p, ctypes.c_void_p, ctypes.c_ulonglong,
                                        ctypes.c_int]
        L.synth_host_sites.restype = ctypes.c_ulonglong
        L.synth_host_sites.argtypes = [ctypes.POINTER(SynthSpec), ctypes.c_void_p, ctypes.c_ulonglong]
        _synth = L
    return _synth


def synth_pileup(seed, sample, genome_len, mean_depth, n_pool_sites, carry, indel_rate, contig, threads=8):
    """One synthetic sample's pileup text as a uint8 array: the bytes snpgpu_synth_pileup_dev writes for the same spec."""
    L = _synth_lib()
    spec = SynthSpec(seed, sample, genome_len, mean_depth, n_pool_sites, carry, indel_rate)
    cap = genome_len * 112 + 4096
    out = np.empty(cap, dtype=np.uint8)
    n = L.synth_host_pileup(ctypes.byref(spec), contig.encode(), out.ctypes.data, cap, threads)
    if n > cap:
        out = np.empty(n, dtype=np.uint8)
        n = L.synth_host_pileup(ctypes.byref(spec), contig.encode(), out.ctypes.data, n, threads)
    return out[:n].copy()


def synth_sample_sites(seed, sample, genome_len, mean_depth, n_pool_sites, carry):
    """1-based positions of the pool sites the sample carries, ascending (uint32)."""
    L = _synth_lib()
    spec = SynthSpec(seed, sample, genome_len, mean_depth, n_pool_sites, carry, 0.0)
    n = L.synth_host_sites(ctypes.byref(spec), None, 0)
    out = np.empty(max(int(n), 1), dtype=np.uint32)
    L.synth_host_sites(ctypes.byref(spec), out.ctypes.data, n)
    return out[:int(n)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        u8p, i32p, i64p, u32p, u64p = (ctypes.POINTER(t) for t in
                                       (ctypes.c_uint8, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32,
                                        ctypes.c_uint64))
        L.oracle_strip.restype = ctypes.c_size_t
        L.oracle_strip.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
        L.oracle_line_report.restype = None
        L.oracle_line_report.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(Params), i32p]
        L.oracle_pileup_consensus.restype = ctypes.c_int
        L.oracle_pileup_consensus.argtypes = [
            ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, i32p, ctypes.c_int,
            i32p, i64p, ctypes.c_size_t, i32p, i64p, ctypes.c_size_t, ctypes.POINTER(Params), ctypes.c_int,
            u8p, u8p, u8p, i64p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
        L.oracle_merge_sites.restype = ctypes.c_size_t
        L.oracle_merge_sites.argtypes = [u64p, u32p, ctypes.c_size_t, u64p, u32p, u32p]
        L.oracle_distance.restype = None
        L.oracle_distance.argtypes = [u8p, ctypes.c_size_t, ctypes.c_size_t, i32p]
        L.oracle_distance_rows.restype = None
        L.oracle_distance_rows.argtypes = [u8p, ctypes.c_size_t, ctypes.c_size_t, i32p, ctypes.c_size_t,
                                           ctypes.c_size_t]
        L.oracle_depth_sum.restype = ctypes.c_int
        L.oracle_depth_sum.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int64), u64p, u64p]
        L.oracle_distance_rows_strided.restype = None
        L.oracle_distance_rows_strided.argtypes = [u8p, ctypes.c_size_t, ctypes.c_size_t, i32p, ctypes.c_size_t,
                                                   ctypes.c_size_t, ctypes.c_size_t]
        _lib = L
    return _lib


def _ptr(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


# ----------------------------------------------------------------------------- per-line (pileup.py)
def strip_bases(bases: bytes) -> bytes:
    """pileup.py:276-325 Record._strip_unwanted_base_patterns."""
    out = ctypes.create_string_buffer(len(bases) + 1)
    n = lib().oracle_strip(bases, len(bases), out)
    return out.raw[:n]


def line_report(line: bytes, params: Params) -> dict:
    """pileup.py:209-274 + 492-590 on one line: tallies, ranking, consensus byte and fail mask."""
    out = np.zeros(13 + 4 * 128, dtype=np.int32)
    lib().oracle_line_report(line, len(line), ctypes.byref(params), _ptr(out, ctypes.c_int32))
    tot, fwd, rev, common = (out[13 + k * 128: 13 + (k + 1) * 128] for k in range(4))
    n_common = int(out[10])
    as_counter = lambda v: {chr(c): int(v[c]) for c in range(128) if v[c]}
    return {
        "status": int(out[0]), "ntok": int(out[1]),
        "pos": (int(out[3]) << 32) | (int(out[2]) & 0xffffffff),
        "raw_depth": (int(out[5]) << 32) | (int(out[4]) & 0xffffffff),
        "ref": chr(out[6]) if out[6] else "", "good_depth": int(out[7]), "fwd_good_depth": int(out[8]),
        "rev_good_depth": int(out[9]),
        "most_common": [chr(c) for c in common[:n_common]] if n_common else None,
        "total": as_counter(tot), "fwd": as_counter(fwd), "rev": as_counter(rev),
        "cons": chr(out[11]), "fail": int(out[12]),
    }


def fail_names(mask: int, params: Params):
    """Filter names in the order pileup.py:556-584 / call_consensus.py:165-168 appends them (None if none)."""
    names = []
    if mask & FAIL_RAWDPTH:
        names.append("RawDpth")
    if mask & FAIL_VARFREQ:
        names.append("VarFreq" + str(int(100 * params.min_cons_freq)))
    if mask & FAIL_DEPTH:
        names.append("Depth" + str(params.min_cons_depth))
    if mask & FAIL_STRDPTH:
        names.append("StrDpth" + str(params.min_cons_strand_depth))
    if mask & FAIL_STRBIAS:
        names.append("StrBias" + str(int(100 * params.min_cons_strand_bias)))
    if mask & FAIL_REGION:
        names.append("Region")
    return names or None


# ----------------------------------------------------------------------------- per-sample driver
def contig_table(names):
    """names: list of bytes -> (concatenated bytes, int32 offsets) as the C entry point wants them."""
    blob = b"".join(names)
    off = np.zeros(len(names) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(n) for n in names])
    return blob, off


def pileup_consensus(text, snp_list, excluded, params: Params, parse_all=False, want_lines=False):
    """call_consensus.py:142-188 for one sample.

    text      bytes / uint8 ndarray holding the pileup file
    snp_list  [(chrom: str, pos: int)] in snplist.txt order (duplicates allowed)
    excluded  iterable of (chrom, pos)
    Returns the consensus row (bytes, len(snp_list)); with want_lines also (cells, fails, positions) of every
    parsed line in file order.  Raises OracleError where the reference raises.
    """
    buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else np.ascontiguousarray(text)
    excluded = list(excluded)
    names = sorted({c for c, _ in snp_list} | {c for c, _ in excluded})
    idx = {c: i for i, c in enumerate(names)}
    blob, off = contig_table([n.encode() for n in names])
    sc = np.array([idx[c] for c, _ in snp_list], dtype=np.int32)
    sp = np.array([p for _, p in snp_list], dtype=np.int64)
    ec = np.array([idx[c] for c, _ in excluded], dtype=np.int32)
    ep = np.array([p for _, p in excluded], dtype=np.int64)
    row = np.zeros(max(len(snp_list), 1), dtype=np.uint8)
    max_lines = 0
    lc = lf = lp = None
    if want_lines:
        max_lines = int(np.count_nonzero(buf == 10) + np.count_nonzero(buf == 13) + 1)
        lc = np.zeros(max_lines, dtype=np.uint8)
        lf = np.zeros(max_lines, dtype=np.uint8)
        lp = np.zeros(max_lines, dtype=np.int64)
    n_parsed = ctypes.c_size_t(0)
    err_line = ctypes.c_size_t(0)
    u8, i32, i64 = ctypes.c_uint8, ctypes.c_int32, ctypes.c_int64
    rc = lib().oracle_pileup_consensus(
        buf.ctypes.data, buf.size, blob, _ptr(off, i32), len(names),
        _ptr(sc, i32), _ptr(sp, i64), len(snp_list), _ptr(ec, i32), _ptr(ep, i64), len(excluded),
        ctypes.byref(params), 1 if parse_all else 0, _ptr(row, u8),
        _ptr(lc, u8) if want_lines else None, _ptr(lf, u8) if want_lines else None,
        _ptr(lp, i64) if want_lines else None, max_lines, ctypes.byref(n_parsed), ctypes.byref(err_line))
    if rc != 0:
        raise OracleError(rc, err_line.value)
    row_b = row[:len(snp_list)].tobytes()
    if want_lines:
        n = n_parsed.value
        return row_b, (lc[:n], lf[:n], lp[:n])
    return row_b


# ----------------------------------------------------------------------------- per-sample consensus VCF
VCF_FORMAT_STR = "GT:SDP:RD:AD:RDF:RDR:ADF:ADR:FT"


def vcf_record_fields(rep: dict, fail_list, failed_snp_gt=".", preserve_ref_case=False):
    """vcf_writer.py:295-379 (_make_vcf_record_from_pileup) on one line_report(): the values of the VCF columns
    REF, ALT, FILTER and of the sample column, as PyVCF3 1.0.3's Writer prints them (vcf/parser.py: missing -> '.',
    empty FILTER list -> 'PASS', lists comma-joined)."""
    ref = rep["ref"]
    upper_ref = ref.upper()
    if not preserve_ref_case:
        ref = upper_ref
    common = rep["most_common"]
    if common is None:
        alt, gt, ad, adf, adr = [], ".", "0", "0", "0"
    else:
        alt = [b for b in common if b != upper_ref]
        if not alt:
            gt, ad, adf, adr = "0", "0", "0", "0"
        else:
            gt = "0" if common[0] == upper_ref else "1"
            ad = ",".join(str(rep["total"].get(b, 0)) for b in alt)
            adf = ",".join(str(rep["fwd"].get(b, 0)) for b in alt)
            adr = ",".join(str(rep["rev"].get(b, 0)) for b in alt)
        if fail_list:
            gt = "." if failed_snp_gt == "." else ("0" if failed_snp_gt == "0" else "1")
    ft = ";".join(fail_list) if fail_list else "PASS"
    sample = ":".join([gt, str(rep["raw_depth"]), str(rep["total"].get(upper_ref, 0)), ad,
                       str(rep["fwd"].get(upper_ref, 0)), str(rep["rev"].get(upper_ref, 0)), adf, adr, ft])
    return ref, (",".join(alt) if alt else "."), ft, sample


def consensus_vcf_body(text: bytes, snp_list, excluded, params: Params, parse_all=False, failed_snp_gt=".",
                       preserve_ref_case=False) -> str:
    """The data lines of the consensus VCF call_consensus.py:161-184 writes: one per pileup line the Reader yields
    (pileup.py:408-429), in file order.  Pure-Python driver over line_report(): small inputs only."""
    wanted = None if parse_all else (set(snp_list) | set(excluded))
    excluded = set(excluded)
    out = []
    for line in text.decode("ascii").splitlines():          # universal newlines, like open() in text mode
        cols = line.rstrip().split()
        if wanted is not None:
            if (cols[0], int(cols[1])) not in wanted:
                continue
        rep = line_report(line.encode(), params)
        if rep["status"]:
            raise OracleError(rep["status"], 0)
        mask = rep["fail"] | (FAIL_REGION if (cols[0], rep["pos"]) in excluded else 0)
        ref, alt, ft, sample = vcf_record_fields(rep, fail_names(mask, params), failed_snp_gt, preserve_ref_case)
        out.append("\t".join([cols[0], str(rep["pos"]), ".", ref, alt, ".", ft, "NS=1", VCF_FORMAT_STR, sample]) + "\n")
    return "".join(out)


# ----------------------------------------------------------------------------- file formats
_ROW_SPLIT = re.compile("\t| +")


def vcf_sites(path):
    """utils.py:1113-1132 convert_vcf_file_to_snp_set, with PyVCF3's Reader restated.

    PyVCF3 1.0.3 ``vcf/parser.py``: the Reader strips every line and drops blank ones, consumes the leading
    ``##`` meta lines and the ``#CHROM`` header line, then for each following line splits
    ``line.rstrip()`` on ``'\\t| +'`` and takes ``CHROM = row[0]``, ``POS = int(row[1])``.
    Returns the ordered list of (chrom, pos) with duplicates removed (a set in the reference).
    """
    seen = {}
    with open(path, "r") as f:
        lines = (ln.strip() for ln in f)
        lines = [ln for ln in lines if ln]
    i = 0
    while i < len(lines) and lines[i].startswith("##"):
        i += 1
    if i >= len(lines):
        raise ValueError("vcf: no header line")     # PyVCF: next() on the exhausted reader inside __init__
    i += 1                                          # the first non-## line is the column header, whatever it is
    for ln in lines[i:]:
        row = _ROW_SPLIT.split(ln.rstrip())
        if len(row) < 8:
            raise IndexError("vcf: fewer than 8 columns")
        seen[(row[0], int(row[1]))] = True
    return list(seen)


def read_snp_list(path):
    """utils.py:1073-1088 read_snp_position_list."""
    out = []
    with open(path, "r") as f:
        for line in f:
            chrom, pos = line.split()[0:2]
            out.append((chrom, int(pos)))
    return out


def merge_sites_text(sample_dirs, vcf_name="var.flt.vcf", max_snps=-1):
    """merge_sites.py:69-131 + utils.py:1056-1070.  Returns (snplist text, filtered sample-dir text)."""
    unsorted = [d for d in (ln.rstrip() for ln in sample_dirs) if d]
    snp_dict, excluded_dirs = {}, set()
    for d in sorted(unsorted):
        vcf = os.path.join(d, vcf_name)
        if not os.path.isfile(vcf) or os.path.getsize(vcf) == 0:
            continue
        name = os.path.basename(os.path.dirname(vcf))
        sites = vcf_sites(vcf)
        if max_snps >= 0 and len(sites) > max_snps:
            excluded_dirs.add(d)
            continue
        for key in sites:
            snp_dict.setdefault(key, []).append(name)
    text = "".join("%s\t%d\t%d\t%s\n" % (k[0], k[1], len(v), "\t".join(v)) for k, v in sorted(snp_dict.items()))
    filt = "".join("%s\n" % d for d in unsorted if d not in excluded_dirs)
    return text, filt


def merge_sites_keys(keys, sample_of):
    """C restatement of the union on packed keys (tests the radix-sort/unique kernel)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    sample_of = np.ascontiguousarray(sample_of, dtype=np.uint32)
    n = keys.size
    uniq = np.zeros(max(n, 1), dtype=np.uint64)
    cnt = np.zeros(max(n, 1), dtype=np.uint32)
    samples = np.zeros(max(n, 1), dtype=np.uint32)
    u = lib().oracle_merge_sites(_ptr(keys, ctypes.c_uint64), _ptr(sample_of, ctypes.c_uint32), n,
                                 _ptr(uniq, ctypes.c_uint64), _ptr(cnt, ctypes.c_uint32),
                                 _ptr(samples, ctypes.c_uint32))
    return uniq[:u], cnt[:u], samples[:n]


def fasta_text(seq_id: str, seq: str) -> str:
    """Bio.SeqIO FastaWriter as called at call_consensus.py:189-192 (description ""): 60-column wrap."""
    lines = [">%s\n" % seq_id]
    for i in range(0, len(seq), 60):
        lines.append(seq[i:i + 60] + "\n")
    return "".join(lines)


def reference_snp_text(reference_fasta_path, snp_list_path):
    """utils.write_reference_snp_file (utils.py:1091-1110) with Biopython's fasta reader / writer restated: one record
    per reference contig in sorted id order, upper(seq[int(pos) - 1]) per snplist line of that contig."""
    with open(snp_list_path) as f:
        position_list = [line.split()[0:2] for line in f]
    records, cur = {}, None
    with open(reference_fasta_path) as f:
        for line in f:
            if line.startswith(">"):
                cur = line[1:].split(None, 1)[0]
                if cur in records:
                    raise ValueError("Duplicate key '%s'" % cur)
                records[cur] = []
            elif cur is not None:
                records[cur].append("".join(line.split()))
    out = []
    for rid in sorted(records):
        seq = "".join(records[rid])
        out.append(fasta_text(rid, "".join(seq[int(pos) - 1].upper() for chrom, pos in position_list if chrom == rid)))
    return "".join(out)


def read_fasta_matrix(path):
    """distance.py:76-84."""
    seqs = {}
    cur = None
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                cur = line.lstrip(">")
                seqs[cur] = ""
            else:
                seqs[cur] += line
    return seqs


def distance_matrix(rows):
    """utils.py:1135-1165 over all pairs: rows = equal-length byte strings -> int32 [n, n]."""
    n = len(rows)
    s = len(rows[0]) if n else 0
    m = np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(n, s).copy() if n and s else np.zeros((n, max(s, 1)), np.uint8)
    d = np.zeros((n, n), dtype=np.int32)
    if n:
        lib().oracle_distance(_ptr(m, ctypes.c_uint8), n, s, _ptr(d, ctypes.c_int32))
    return d


def distance_matrix_threads(rows, threads=1):
    """distance_matrix() with the rows of the upper triangle dealt to `threads` host threads (the C call drops the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    n = len(rows)
    s = len(rows[0]) if n else 0
    if n == 0 or s == 0 or threads <= 1:
        return distance_matrix(rows)
    m = np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(n, s).copy()
    d = np.zeros((n, n), dtype=np.int32)
    L = lib()
    # row i of the triangle costs n - i pairs: interleave the rows so that every thread gets the same share

    def work(t):
        L.oracle_distance_rows_strided(_ptr(m, ctypes.c_uint8), n, s, _ptr(d, ctypes.c_int32), t, n, threads)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, range(threads)))
    return d


def hot_path_texts(names, texts, site_lists, params, parse_all=False, threads=1, rows=None):
    """The path's three files for samples given in sorted order, restated on the CPU: (snplist.txt, snpma.fasta,
    snp_distance_matrix.tsv) texts.  texts: per sample the pileup bytes (or None when `rows` already holds the
    consensus rows); site_lists: per sample [(chrom, pos)].  merge_sites.py:94-116 + utils.py:1056-1070,
    call_consensus.py:161-192, snp_matrix.py:114-119, distance.py:90-115."""
    from concurrent.futures import ThreadPoolExecutor
    chroms = sorted({c for s in site_lists for c, _ in s})
    rank = {c: i for i, c in enumerate(chroms)}
    per = [list(dict.fromkeys(s)) for s in site_lists]
    keys = np.array([(rank[c] << 32) | p for s in per for c, p in s], dtype=np.uint64)
    samp = np.array([i for i, s in enumerate(per) for _ in s], dtype=np.uint32)
    uniq, cnt, samples = merge_sites_keys(keys, samp)
    out, o = [], 0
    sl = samples.tolist()
    for k, c in zip(uniq.tolist(), cnt.tolist()):
        out.append("%s\t%d\t%d\t%s\n" % (chroms[k >> 32], k & 0xffffffff, c, "\t".join(names[i] for i in sl[o:o + c])))
        o += c
    snplist = "".join(out)
    snps = [(chroms[k >> 32], k & 0xffffffff) for k in uniq.tolist()]
    if rows is None:
        with ThreadPoolExecutor(max_workers=max(threads, 1)) as ex:
            rows = list(ex.map(lambda t: pileup_consensus(t, snps, [], params, parse_all=parse_all), texts))
    snpma = "".join(fasta_text(n, r.decode("ascii")) for n, r in zip(names, rows))
    order = sorted(range(len(names)), key=lambda i: names[i])
    d = distance_matrix_threads([rows[i] for i in order], threads) if names else np.zeros((0, 0), np.int32)
    ids = [names[i] for i in order]
    mat = ["\t%s\n" % "\t".join(ids)]
    for i, a in enumerate(ids):
        mat.append("%s\t%s\n" % (a, "\t".join(map(str, d[i].tolist()))))
    return snplist, snpma, "".join(mat), rows


def depth_sum(text: bytes):
    """collect_metrics.py:322-329: (sum of int(line.split()[3]) over the lines, lines that contributed)."""
    total, lines, off = ctypes.c_int64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
    rc = lib().oracle_depth_sum(bytes(text), len(text), ctypes.byref(total), ctypes.byref(lines), ctypes.byref(off))
    if rc:
        raise OracleError(rc, off.value)
    return total.value, lines.value


def mean_pileup_depth_text(text: bytes, reference_length: int) -> str:
    """collect_metrics.py:333-338: "%.2f" of depth_sum / reference_length, "" when either is not positive."""
    total, _ = depth_sum(text)
    if total > 0 and reference_length > 0:
        return "%.2f" % (float(total) / float(reference_length))
    return ""


def distance_texts(seqs):
    """distance.py:90-115: (pairwise TSV text, matrix TSV text) from {id: sequence}."""
    ids = sorted(seqs)
    for a, b in itertools.combinations(ids, 2):
        if len(seqs[b]) < len(seqs[a]):
            raise IndexError("string index out of range")
    # the reference walks range(len(seq1)) for each sorted pair (a, b): compare over len(a)
    lens = [len(seqs[i]) for i in ids]
    if len(set(lens)) <= 1:
        d = distance_matrix([seqs[i].encode() for i in ids]) if ids else np.zeros((0, 0), np.int32)
    else:
        d = np.zeros((len(ids), len(ids)), np.int32)
        for (i, a), (j, b) in itertools.combinations(enumerate(ids), 2):
            la = len(seqs[a])
            d[i, j] = d[j, i] = distance_matrix([seqs[a].encode(), seqs[b][:la].encode()])[0, 1]
    pair = ["Seq1\tSeq2\tDistance\n"]
    for i, a in enumerate(ids):
        for j, b in enumerate(ids):
            pair.append("%s\t%s\t%i\n" % (a, b, d[i, j]))
    mat = ["\t%s\n" % "\t".join(ids)]
    for i, a in enumerate(ids):
        mat.append("%s\t%s\n" % (a, "\t".join(str(int(x)) for x in d[i])))
    return "".join(pair), "".join(mat)


# ---------------------------------------------------------------------------------------------------------------
# filter_regions: the arithmetic and the file split, restated in plain Python
# ---------------------------------------------------------------------------------------------------------------
def merge_regions(regions):
    """utils.py:1168-1282: coalesce overlapping, contained and adjacent (start, end) regions."""
    if len(regions) == 0:
        return regions
    regions = sorted(regions)
    merged = [regions[0]]
    for start, end in regions[1:]:
        last_start, last_end = merged[-1]
        if start >= last_start and end <= last_end:
            continue
        if start <= last_end + 1 and end > last_end:
            merged[-1] = (last_start, end)
        else:
            merged.append((start, end))
    return merged


def in_region(pos, regions):
    """utils.py:1285-1318."""
    return any(start <= pos <= end for start, end in regions)


def find_dense_regions(max_allowed_snps, window_size, snps):
    """filter_regions.py:17-71 over a sorted list of positions."""
    out = []
    for idx, pos_start in enumerate(snps):
        if idx + max_allowed_snps < len(snps):
            pos_end = snps[idx + max_allowed_snps]
            if pos_start + window_size - 1 >= pos_end:
                out.append((pos_start, pos_end))
    return merge_regions(out)


def collect_dense_regions(records, bad_regions, contig_length, edge_length, max_snps_list, window_size_list):
    """filter_regions.py:386-428: records = [(chrom, pos)] of one sample; bad_regions (chrom -> list) is extended."""
    snp_dict = {}
    for chrom, pos in records:
        snp_dict.setdefault(chrom, []).append(pos)
    for contig, snp_list in snp_dict.items():
        if contig not in bad_regions:
            length = contig_length.get(contig, sys.maxsize)
            if length <= edge_length * 2:
                bad_regions[contig] = [(0, length)]
            else:
                bad_regions[contig] = [(0, edge_length), (length - edge_length, length)]
        sorted_snps = sorted(snp_list)
        for max_allowed, window in zip(max_snps_list, window_size_list):
            bad_regions[contig].extend(find_dense_regions(max_allowed, window, sorted_snps))


def filter_regions_flags(samples, contig_length, edge_length=500, window_size_list=(1000,), max_snps_list=(3,), mode="all",
                         outgroup=()):
    """removed[s][r] for record r of sample s (None for an outgroup sample: its file is copied, filter_regions.py:431-457).
    samples: [(sample_id, [(chrom, pos) in file order])].  filter_regions.py:203-303 (mode "all") / :306-383 ("each")."""
    flags = []
    if mode == "all":
        bad = {}
        for sid, records in samples:
            if sid not in outgroup:
                collect_dense_regions(records, bad, contig_length, edge_length, max_snps_list, window_size_list)
        bad = {c: merge_regions(r) for c, r in bad.items()}
        for sid, records in samples:
            flags.append(None if sid in outgroup else [in_region(p, bad[c]) for c, p in records])
    else:
        for sid, records in samples:
            if sid in outgroup:
                flags.append(None)
                continue
            bad = {}
            collect_dense_regions(records, bad, contig_length, edge_length, max_snps_list, window_size_list)
            bad = {c: merge_regions(r) for c, r in bad.items()}
            flags.append([in_region(p, bad[c]) for c, p in records])
    return flags


def vcf_split_texts(vcf_text, flags):
    """(preserved, removed) file texts of filter_regions.py:460-520 for a VarScan-style VCF.  PyVCF3's Writer puts the
    template's header back as: the plain ##key=value lines, then ##INFO, ##FORMAT, ##FILTER, ##ALT, ##contig, then the column
    line; records are echoed (PyVCF3 round-trips the VarScan records of the reference's data sets character for character:
    pinned by the lambda / agona golden files).  flags None: the outgroup form (header only in the removed file)."""
    lines = [ln for ln in vcf_text.split("\n") if ln.strip()]
    meta = [ln for ln in lines if ln.startswith("##")]
    rest = [ln for ln in lines if not ln.startswith("##")]
    kinds = ("##INFO=", "##FORMAT=", "##FILTER=", "##ALT=", "##contig=")
    head = [ln for ln in meta if not ln.startswith(kinds)]
    for k in kinds:
        head += [ln for ln in meta if ln.startswith(k)]
    head = "".join(ln + "\n" for ln in head + rest[:1])
    records = rest[1:]
    if flags is None:
        return vcf_text, head
    assert len(flags) == len(records)
    keep = "".join(ln + "\n" for ln, f in zip(records, flags) if not f)
    drop = "".join(ln + "\n" for ln, f in zip(records, flags) if f)
    return head + keep, head + drop
