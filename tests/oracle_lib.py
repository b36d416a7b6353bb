"""ctypes wrapper over oracle/liboracle.so — TEST INFRASTRUCTURE (the checker, never the product)."""
import ctypes as C
import os
import subprocess

from stitch_b200 import _abi, _lib
from stitch_b200._abi import StitchChain, StitchContig, StitchOp, StitchOpts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")

ORACLE_RESULTS = {"n_reads": "oracle_result_n_reads", "read": "oracle_result_read",
                  "chains": "oracle_result_chains", "ops": "oracle_result_ops", "free": "oracle_result_free"}

_o = None


def build():
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("stitch_oracle.cpp", "oracle_capi.cpp", "stitch_oracle.hpp")]
    if os.path.exists(ORACLE_SO) and all(os.path.getmtime(ORACLE_SO) >= os.path.getmtime(s) for s in srcs if os.path.exists(s)):
        return
    try:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"])
    except (subprocess.CalledProcessError, OSError):
        if not os.path.exists(ORACLE_SO):   # (a box that cannot rebuild uses the library shipped with the tree)
            raise


def lib():
    global _o
    if _o is not None:
        return _o
    if not os.path.exists(ORACLE_SO):
        build()
    o = C.CDLL(ORACLE_SO)
    o.oracle_last_error.restype = C.c_char_p
    o.oracle_sca.restype = C.c_int
    o.oracle_sca.argtypes = [C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_char_p, C.c_int64, C.c_char_p,
                             C.c_int64, C.POINTER(C.c_void_p)]
    o.oracle_mca.restype = C.c_int
    o.oracle_mca.argtypes = [C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_char_p),
                             C.c_char_p, C.c_char_p, C.POINTER(C.c_int32), C.c_char_p, C.c_int64,
                             C.c_char_p, C.POINTER(C.c_void_p)]
    o.oracle_split_at_y.restype = C.c_int
    o.oracle_split_at_y.argtypes = [C.POINTER(StitchChain), C.POINTER(StitchOp), C.c_int, C.c_int64,
                                    C.POINTER(C.c_void_p)]
    o.oracle_cigar.restype = C.c_int
    o.oracle_cigar.argtypes = [C.POINTER(StitchChain), C.POINTER(StitchOp), C.c_char_p, C.c_size_t]
    o.oracle_aligner_create.restype = C.c_int
    o.oracle_aligner_create.argtypes = [C.POINTER(StitchOpts), C.POINTER(StitchContig), C.c_uint32,
                                        C.POINTER(C.c_void_p)]
    o.oracle_aligner_destroy.restype = None
    o.oracle_aligner_destroy.argtypes = [C.c_void_p]
    o.oracle_aligner_batch.restype = C.c_int
    o.oracle_aligner_batch.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_uint32, C.c_void_p,
                                       C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    _lib.declare_results_api(o, ORACLE_RESULTS)
    o.oracle_result_cigar.restype = C.c_char_p
    o.oracle_result_cigar.argtypes = [C.c_void_p, C.c_uint64]
    o.oracle_result_cells.restype = C.c_uint64
    o.oracle_result_cells.argtypes = [C.c_void_p]
    o.oracle_result_fills.restype = C.c_uint64
    o.oracle_result_fills.argtypes = [C.c_void_p]
    o.oracle_result_seconds.restype = C.c_double
    o.oracle_result_seconds.argtypes = [C.c_void_p]
    o.oracle_set_checker_layout.restype = None
    o.oracle_set_checker_layout.argtypes = [C.c_int]
    _o = o
    return o


def set_checker_layout(colmajor: bool):
    """Traceback matrix layout of the oracle: False = the reference's (i * cols + j; what the CPU baseline is timed on),
    True = column-major (same values, several times faster on full-length reads; parity checks only)."""
    lib().oracle_set_checker_layout(int(bool(colmajor)))


class OracleError(RuntimeError):
    pass


def _take(handle, with_cigars=False):
    o = lib()
    try:
        per_read = _lib.read_results(o, ORACLE_RESULTS, handle)
        info = {"cells": o.oracle_result_cells(handle), "fills": o.oracle_result_fills(handle),
                "seconds": o.oracle_result_seconds(handle)}
        if with_cigars:
            k = 0
            for lst in per_read:
                for a in lst:
                    a.oracle_cigar = o.oracle_result_cigar(handle, k).decode()
                    k += 1
        return per_read, info
    finally:
        o.oracle_result_free(handle)


def sca(mode, x, y, match=1, mismatch=-1, gap_open=-5, gap_extend=-1, jump=-10, circular=False):
    """SingleContigAligner::{local,querylocal,targetlocal,global} (mode 0..3)."""
    o = lib()
    sc = (C.c_int32 * 5)(match, mismatch, gap_open, gap_extend, jump)
    h = C.c_void_p()
    rc = o.oracle_sca(mode, sc, int(circular), bytes(x), len(x), bytes(y), len(y), C.byref(h))
    if rc != 0:
        raise OracleError(o.oracle_last_error().decode())
    per_read, _ = _take(h, with_cigars=True)
    return per_read[0][0]


def mca(contigs, scoring, y, subset=None):
    """MultiContigAligner::custom over explicit contig-strands.
    contigs: list of dict(name, fwd(bool), seq, circular?)."""
    o = lib()
    n = len(contigs)
    seqs = (C.c_char_p * n)(*[bytes(c["seq"]) for c in contigs])
    lens = (C.c_int64 * n)(*[len(c["seq"]) for c in contigs])
    names = (C.c_char_p * n)(*[c["name"].encode() for c in contigs])
    fwd = bytes(int(bool(c["fwd"])) for c in contigs)
    circ = bytes(int(bool(c.get("circular", False))) for c in contigs)
    keys = ["match", "mismatch", "gap_open", "gap_extend", "jump_same", "jump_opp", "jump_inter",
            "xclip_prefix", "xclip_suffix", "yclip_prefix", "yclip_suffix"]
    sc = (C.c_int32 * 11)(*[scoring[k] for k in keys])
    sub = None if subset is None else bytes(int(bool(b)) for b in subset)
    h = C.c_void_p()
    rc = o.oracle_mca(n, seqs, lens, names, fwd, circ, sc, bytes(y), len(y), sub, C.byref(h))
    if rc != 0:
        raise OracleError(o.oracle_last_error().decode())
    per_read, _ = _take(h, with_cigars=True)
    return per_read[0][0]


def _chain_struct(a):
    ops = (StitchOp * max(1, len(a["ops"])))()
    for k, (kind, x, y) in enumerate(a["ops"]):
        ops[k] = StitchOp(kind, x, y)
    ch = StitchChain(score=a["score"], xstart=a["xstart"], xend=a["xend"], ystart=a["ystart"], yend=a["yend"],
                     xlen=a["xlen"], ylen=a["ylen"], start_contig_idx=a["start_contig_idx"],
                     end_contig_idx=a["end_contig_idx"], length=a["length"], n_ops=len(a["ops"]), ops_offset=0)
    return ch, ops


def split_at_y(alignment: dict, y_pivot: int):
    o = lib()
    ch, ops = _chain_struct(alignment)
    h = C.c_void_p()
    rc = o.oracle_split_at_y(C.byref(ch), ops, alignment["mode"], y_pivot, C.byref(h))
    if rc != 0:
        raise OracleError(o.oracle_last_error().decode())
    per_read, _ = _take(h, with_cigars=True)
    return per_read[0][0]


def cigar_of(a) -> str:
    """Alignment::cigar() computed by the oracle for any Alignment (oracle's or product's)."""
    o = lib()
    ch, ops = _chain_struct(dict(score=a.score, xstart=a.xstart, xend=a.xend, ystart=a.ystart, yend=a.yend,
                                 xlen=a.xlen, ylen=a.ylen, start_contig_idx=a.start_contig_idx,
                                 end_contig_idx=a.end_contig_idx, length=a.length, ops=a.ops))
    buf = C.create_string_buffer(16 * (len(a.ops) + 4) + 64)
    rc = o.oracle_cigar(C.byref(ch), ops, buf, len(buf))
    if rc != 0:
        raise OracleError(o.oracle_last_error().decode())
    return buf.value.decode()


class OracleAligners:
    """Builder::build_aligners + Aligners::align on the CPU oracle."""

    def __init__(self, opts: StitchOpts, contigs):
        o = lib()
        arr, self._keep = _abi.make_contigs(contigs)
        self.n_strands = len(contigs) * (2 if opts.double_strand else 1)
        h = C.c_void_p()
        rc = o.oracle_aligner_create(C.byref(opts), arr, len(contigs), C.byref(h))
        if rc != 0:
            raise OracleError(o.oracle_last_error().decode())
        self._h = h

    def batch(self, reads, subsets=None, raw=False, threads=1):
        o = lib()
        buf, offs = _abi.pack_reads(reads)
        words, stride = None, 0
        if subsets is not None:
            stride = (self.n_strands + 31) // 32
            words = (C.c_uint32 * (stride * len(reads)))()
            for r, sub in enumerate(subsets):
                for c in (sub or ()):
                    words[r * stride + c // 32] |= 1 << (c % 32)
        h = C.c_void_p()
        rc = o.oracle_aligner_batch(self._h, buf, offs, len(reads), words, stride, int(raw), int(threads), C.byref(h))
        if rc != 0:
            raise OracleError(o.oracle_last_error().decode())
        return _take(h, with_cigars=True)

    def close(self):
        if self._h:
            lib().oracle_aligner_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
