"""CPU tier: the product library loads, exports every symbol include/stitch_b200.h declares, and
fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    from stitch_b200 import _lib
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "stitch_b200.h")).read()
    declared = set(re.findall(r"\b(stitch_[a-z0-9_]+)\s*\(", header))
    from stitch_b200 import _lib
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.stitch_abi_version() == 3


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import stitch_b200
    with pytest.raises(stitch_b200.StitchError) as e:
        stitch_b200.Builder().build_aligners([stitch_b200.TargetSeq("c", b"ACGT")])
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_references_oracle():
    """The product tree must not import, include or load anything under oracle/ or tests/."""
    pkg = os.path.join(ROOT, "stitch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".hpp", ".cu", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                for line in text.split("\n"):
                    code = line.split("//")[0].split("#")[0] if not line.lstrip().startswith("#include") else line
                    assert "oracle_lib" not in code and "liboracle" not in code and "emul_backend" not in code, (f, line)
                    assert not re.search(r'#include\s+"[^"]*oracle', code), (f, line)


def test_invalid_options_are_rejected(lib):
    """Sign checks of Scoring::with_jump_scores (scoring.rs:36-75) and Options::clipping (mod.rs:129)
    happen before any device work."""
    from stitch_b200._abi import make_contigs, make_opts
    arr, keep = make_contigs([("c", b"ACGT")])
    for kw in (dict(gap_open=1), dict(gap_extend=2), dict(default_jump_score=3), dict(mode=4)):
        h = C.c_void_p()
        rc = lib.stitch_create(C.byref(make_opts(**kw)), arr, 1, 0, C.byref(h))
        assert rc == -1, kw
        assert lib.stitch_last_error(None)
