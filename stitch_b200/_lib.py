"""Loads the CUDA/C-ABI library (stitch_b200/libstitch_b200.so) and declares its signatures.

There is no fallback: if the library has not been built (`python -c "import __graft_entry__ as g;
g.build()"`) importing any compute entry point fails loudly.
"""
import ctypes as C
import os

from ._abi import StitchChain, StitchContig, StitchOp, StitchOpts, StitchStats

_HERE = os.path.dirname(os.path.abspath(__file__))
# (STITCH_B200_LIB selects another build of the same sources, e.g. a -DSTITCH_PACK_WARPS=12 variant; the CLI honours it too)
LIB_PATH = os.environ.get("STITCH_B200_LIB") or os.path.join(_HERE, "libstitch_b200.so")

# every symbol include/stitch_b200.h declares
EXPORTED_SYMBOLS = [
    "stitch_create", "stitch_align_batch", "stitch_custom_batch", "stitch_custom_batch_device",
    "stitch_results_n_reads", "stitch_results_read", "stitch_results_chains", "stitch_results_ops",
    "stitch_free_results", "stitch_get_stats", "stitch_set_max_inflight", "stitch_destroy",
    "stitch_last_error", "stitch_abi_version", "stitch_measure_int32_peak", "stitch_format_sam", "stitch_free_text",
    "stitch_results_from_chains", "stitch_results_prealign", "stitch_prealign_batch",
]

_lib = None


def declare_results_api(lib, prefix_map):
    """Shared by the product library and (in tests) the oracle, whose accessors have the same shape."""
    n_reads = getattr(lib, prefix_map["n_reads"])
    n_reads.restype = C.c_uint32
    n_reads.argtypes = [C.c_void_p]
    read = getattr(lib, prefix_map["read"])
    read.restype = None
    read.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    chains = getattr(lib, prefix_map["chains"])
    chains.restype = C.POINTER(StitchChain)
    chains.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    ops = getattr(lib, prefix_map["ops"])
    ops.restype = C.POINTER(StitchOp)
    ops.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    free = getattr(lib, prefix_map["free"])
    free.restype = None
    free.argtypes = [C.c_void_p]


def results_from_chains(lib, prefix, chains):
    """A results handle (one read) holding `chains` (Alignment objects); free it with <prefix>free_results."""
    n_ops = sum(len(a.ops) for a in chains)
    ch = (StitchChain * max(1, len(chains)))()
    ops = (StitchOp * max(1, n_ops))()
    k = 0
    for c, a in enumerate(chains):
        ch[c] = StitchChain(score=a.score, xstart=a.xstart, xend=a.xend, ystart=a.ystart, yend=a.yend, xlen=a.xlen, ylen=a.ylen,
                            start_contig_idx=a.start_contig_idx, end_contig_idx=a.end_contig_idx, length=a.length,
                            n_ops=len(a.ops), ops_offset=k)
        for kind, x, y in a.ops:
            ops[k] = StitchOp(kind, x, y)
            k += 1
    f = getattr(lib, prefix + "results_from_chains")
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    out = C.c_void_p()
    rc = f(ch, len(chains), ops, n_ops, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"results_from_chains failed ({rc})")
    return out


def declare_sam_api(lib, prefix):
    f = getattr(lib, prefix + "format_sam")
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int32,
                  C.c_void_p, C.POINTER(C.c_void_p)]
    g = getattr(lib, prefix + "free_text")
    g.restype = None
    g.argtypes = [C.c_void_p]


def format_sam(lib, prefix, ctx, results_handle, read, header, bases, quals=None, pre_align_score=None, sam_opts=None):
    """SAM text lines of one read of a results handle (stitch_format_sam)."""
    from ._abi import StitchSamOpts
    bases = bytes(bases)
    bbuf = (C.c_uint8 * max(1, len(bases))).from_buffer_copy(bases if bases else b"\0")
    qbuf = None
    if quals is not None:
        quals = bytes(quals)
        qbuf = (C.c_uint8 * max(1, len(quals))).from_buffer_copy(quals if quals else b"\0")
    so = None
    if sam_opts:
        so = StitchSamOpts(soft_clip=int(bool(sam_opts.get("soft_clip", False))), use_eq_and_x=int(bool(sam_opts.get("use_eq_and_x", False))),
                           pick_primary=int(sam_opts.get("pick_primary", 0)), filter_secondary=int(bool(sam_opts.get("filter_secondary", False))),
                           filter_secondary_pct=float(sam_opts.get("filter_secondary_pct", 10.0)))
    out = C.c_void_p()
    hdr = header.encode() if isinstance(header, str) else bytes(header)
    rc = getattr(lib, prefix + "format_sam")(ctx, results_handle, read, hdr, bbuf, qbuf, len(bases), int(pre_align_score is not None),
                                             int(pre_align_score or 0), C.byref(so) if so else None, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"format_sam failed ({rc}): {getattr(lib, prefix + 'last_error')(ctx).decode()}")
    try:
        text = C.string_at(out).decode()
        return text.split("\n") if text else []   # (every record can be filtered out by filter_secondary)
    finally:
        getattr(lib, prefix + "free_text")(out)


def read_results(lib, prefix_map, handle):
    """-> list (per read) of lists of Alignment."""
    from .alignment import Alignment
    n_reads = getattr(lib, prefix_map["n_reads"])(handle)
    n_ch, n_op = C.c_uint64(), C.c_uint64()
    chains = getattr(lib, prefix_map["chains"])(handle, C.byref(n_ch))
    ops = getattr(lib, prefix_map["ops"])(handle, C.byref(n_op))
    out = []
    first, count = C.c_uint64(), C.c_uint32()
    for r in range(n_reads):
        getattr(lib, prefix_map["read"])(handle, r, C.byref(first), C.byref(count))
        lst = []
        for k in range(first.value, first.value + count.value):
            c = chains[k]
            o = [(ops[t].kind, ops[t].a, ops[t].b) for t in range(c.ops_offset, c.ops_offset + c.n_ops)]
            lst.append(Alignment(c.score, c.xstart, c.xend, c.ystart, c.yend, c.xlen, c.ylen,
                                 c.start_contig_idx, c.end_contig_idx, c.length, o))
        out.append(lst)
    return out


def prealign_batch(lib, prefix, ctx, reads, n_strands):
    """stitch_prealign_batch: -> (selected contig-strands per read, best score per read)."""
    from ._abi import pack_reads
    f = getattr(lib, prefix + "prealign_batch")
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_int32)]
    buf, offs = pack_reads(reads)
    stride = (n_strands + 31) // 32
    words = (C.c_uint32 * max(1, stride * len(reads)))()
    best = (C.c_int32 * max(1, len(reads)))()
    rc = f(ctx, buf, offs, len(reads), words, stride, best)
    if rc != 0:
        raise RuntimeError(f"prealign_batch failed ({rc}): {getattr(lib, prefix + 'last_error')(ctx).decode()}")
    sel = []
    for r in range(len(reads)):
        row = []
        for w in range(stride):
            bits = words[r * stride + w]
            while bits:
                low = bits & -bits
                row.append(w * 32 + low.bit_length() - 1)
                bits ^= low
        sel.append(row)
    return sel, [int(best[r]) for r in range(len(reads))]


def read_prealign(lib, prefix, handle, n_reads):
    """The Option<i32> pre-alignment score of every read of a results handle (None = no score)."""
    f = getattr(lib, prefix + "results_prealign")
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_int32)]
    out = []
    v = C.c_int32()
    for r in range(n_reads):
        out.append(int(v.value) if f(handle, r, C.byref(v)) else None)
    return out


PRODUCT_RESULTS = {"n_reads": "stitch_results_n_reads", "read": "stitch_results_read",
                   "chains": "stitch_results_chains", "ops": "stitch_results_ops", "free": "stitch_free_results"}


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.stitch_create.restype = C.c_int
    lib.stitch_create.argtypes = [C.POINTER(StitchOpts), C.POINTER(StitchContig), C.c_uint32, C.c_int,
                                  C.POINTER(C.c_void_p)]
    for name in ("stitch_align_batch", "stitch_custom_batch"):
        f = getattr(lib, name)
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_uint32, C.c_void_p,
                      C.c_uint32, C.POINTER(C.c_void_p)]
    lib.stitch_custom_batch_device.restype = C.c_int
    lib.stitch_custom_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_uint32,
                                               C.POINTER(C.c_void_p)]
    declare_results_api(lib, PRODUCT_RESULTS)
    declare_sam_api(lib, "stitch_")
    lib.stitch_get_stats.restype = C.c_int
    lib.stitch_get_stats.argtypes = [C.c_void_p, C.POINTER(StitchStats)]
    lib.stitch_set_max_inflight.restype = C.c_int
    lib.stitch_set_max_inflight.argtypes = [C.c_void_p, C.c_uint32]
    lib.stitch_destroy.restype = None
    lib.stitch_destroy.argtypes = [C.c_void_p]
    lib.stitch_last_error.restype = C.c_char_p
    lib.stitch_last_error.argtypes = [C.c_void_p]
    lib.stitch_measure_int32_peak.restype = C.c_int
    lib.stitch_measure_int32_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.stitch_abi_version.restype = C.c_uint32
    lib.stitch_abi_version.argtypes = []
    _lib = lib
    return lib
