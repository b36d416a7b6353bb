"""Hand-derived known-answer records of the SAM record layer (tests/golden/sam_kats.json, written by
tests/golden/make_sam_kats.py from the reference source, one case per quirk): both the product's formatter
(stitch_b200/csrc/host_sam.hpp behind stitch_format_sam, here through the CPU emulator library, and through the CUDA
library in the gpu tier) and the checker (oracle/sam_oracle.py) must reproduce them."""
import ctypes as C
import json
import os
import sys

import pytest

from stitch_b200 import _abi, _lib
from stitch_b200._abi import make_opts
from stitch_b200.alignment import Alignment
from test_emul_parity import emul_lib  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sam_oracle  # noqa: E402

CASES = json.load(open(os.path.join(ROOT, "tests", "golden", "sam_kats.json")))


def chains_of(case):
    return [Alignment(c["score"], c["xstart"], c["xend"], c["ystart"], c["yend"], c["xlen"], c["ylen"], c["start_contig_idx"],
                      c["end_contig_idx"], c["length"], [tuple(o) for o in c["ops"]]) for c in case["chains"]]


def product_lines(lib, prefix, create, destroy, case):
    """stitch_results_from_chains + stitch_format_sam on a context built over the case's targets."""
    m, x, o, e = case["scoring"]
    opts = make_opts(match_score=m, mismatch_score=x, gap_open=o, gap_extend=e, double_strand=case["double_strand"])
    arr, keep = _abi.make_contigs([(n, b"A" * l) for n, l in case["targets"]])
    h = C.c_void_p()
    assert create(C.byref(opts), arr, len(case["targets"]), 0, C.byref(h)) == 0
    res = _lib.results_from_chains(lib, prefix, chains_of(case))
    try:
        return _lib.format_sam(lib, prefix, h, res, 0, case["header"], case["bases"].encode(),
                               None if case["quals"] is None else case["quals"].encode(), case["pre_align_score"], case["sam_opts"])
    finally:
        getattr(lib, prefix + "free_results")(res)
        destroy(h)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_sam_kat_checker(case):
    got = sam_oracle.format_sam(case["header"], case["bases"].encode(), None if case["quals"] is None else case["quals"].encode(),
                                chains_of(case), [tuple(t) for t in case["targets"]], tuple(case["scoring"]),
                                pre_alignment_score=case["pre_align_score"], **case["sam_opts"])
    assert got == case["expect"]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_sam_kat_product_host_code(emul_lib, case):
    e = emul_lib.lib(8)
    assert product_lines(e, "emul_", e.emul_create, e.emul_destroy, case) == case["expect"]


@pytest.mark.gpu
def test_sam_kats_on_gpu():
    lib = _lib.load()
    for case in CASES:
        assert product_lines(lib, "stitch_", lib.stitch_create, lib.stitch_destroy, case) == case["expect"], case["name"]
