"""CPU tier: quiet tiles (stitch_b200/csrc/dp_packed.h, PkQuiet).  The bulk pass of the packed fill skips warp
tiles whose cells are provably in the closed form "jump + substitution score" and re-materialises their
state when it is needed again.  Run through the CPU emulator (skipped tiles leave POISON in the state
arrays, so any read of stale state shows up as a parity failure) against the oracle, with a checkpoint
spacing large enough for tiles to go quiet, and check that tiles were in fact skipped."""
import ctypes as C
import os
import random

import pytest

import gen
from test_emul_parity import compare, emul_lib, run_both  # noqa: F401


def counters(emul_lib, strip):
    out = (C.c_ulonglong * 3)()
    emul_lib.lib(strip).emul_quiet_counters(out)
    return list(out)


@pytest.fixture
def wide_checkpoints(monkeypatch):
    monkeypatch.setenv("EMUL_K", "48")
    monkeypatch.setenv("EMUL_QUIET", "1")


@pytest.mark.parametrize("strip", [1, 2, 8])
@pytest.mark.parametrize("block", range(4))
def test_quiet_multi_tile(oracle, emul_lib, wide_checkpoints, strip, block):
    tile = 32 * strip
    before = counters(emul_lib, strip)
    for seed in range(block * 6, block * 6 + 6):
        rng = random.Random(77000 + seed * 13 + strip)
        alphabet = [b"ACGT", b"ACGT", b"ACG", b"ACGTN"][seed % 4]
        lens = [rng.randint(3 * tile, 9 * tile) for _ in range(rng.randint(1, 4))] + [tile + 1, 2 * tile][: seed % 3]
        contigs = [gen.rand_seq(rng, l, alphabet) for l in lens]
        nread = rng.randint(120, 420 if strip < 8 else 300)
        reads = [gen.chimeric_read(rng, contigs, nread, rng.randint(1, 4), strands=seed % 2 == 0,
                                   wrap=rng.random() < 0.5, noise=rng.random() < 0.8, alphabet=alphabet) for _ in range(2)]
        reads = [r if r else b"A" for r in reads]
        kw = gen.fuzz_opts_packed(seed + 50 * block, strip) if seed % 3 else dict(
            mode=rng.randint(0, 3), double_strand=seed % 2 == 0, circular=rng.random() < 0.5)
        if strip < 8 and seed % 3 == 0:
            kw = gen.fuzz_opts_packed(seed + 50 * block, strip)
        got, exp = run_both(oracle, emul_lib, kw, contigs, reads, strip, raw=(seed % 2 == 0))
        compare(got, exp, f"seed {seed} strip {strip} {kw}")
    after = counters(emul_lib, strip)
    assert after[0] > before[0]


def test_quiet_tiles_are_skipped(oracle, emul_lib, wide_checkpoints):
    """Reference CLI defaults on chimeric ONT-like reads: a good share of the tile-columns is skipped, and the
    result is still the oracle's."""
    rng = random.Random(5)
    contigs = [gen.rand_seq(rng, rng.randint(2600, 3400)) for _ in range(3)]
    reads = [gen.chimeric_read(rng, contigs, 500, 3, strands=True, wrap=True) for _ in range(2)]
    before = counters(emul_lib, 8)
    for kw in (dict(double_strand=True, circular=True), dict(mode=1, double_strand=True), dict(mode=3, double_strand=True)):
        got, exp = run_both(oracle, emul_lib, kw, contigs, reads, 8, raw=False)
        compare(got, exp, str(kw))
    after = counters(emul_lib, 8)
    tiles, skipped = after[0] - before[0], after[1] - before[1]
    assert tiles > 0 and skipped > 0.25 * tiles, (tiles, skipped)


def test_quiet_off_is_identical(oracle, emul_lib, monkeypatch):
    monkeypatch.setenv("EMUL_K", "48")
    monkeypatch.setenv("EMUL_QUIET", "0")
    rng = random.Random(6)
    contigs = [gen.rand_seq(rng, 1500) for _ in range(2)]
    reads = [gen.chimeric_read(rng, contigs, 300, 3, strands=True, wrap=True)]
    before = counters(emul_lib, 8)
    got, exp = run_both(oracle, emul_lib, dict(double_strand=True, circular=True), contigs, reads, 8, raw=False)
    compare(got, exp, "quiet off")
    assert counters(emul_lib, 8)[1] == before[1]


@pytest.mark.parametrize("case", range(4))
def test_quiet_special_tiles(oracle, emul_lib, wide_checkpoints, case):
    """First / last tiles of contigs and of warp chunks under the closed form: contig lengths around the tile size (row m
    alone in the last tile, m a multiple of the tile), circular wrap into row 1, all four modes, tie-prone alphabet."""
    rng = random.Random(88000 + case)
    kw = [dict(double_strand=True, circular=True), dict(mode=3, circular=True), dict(mode=1, double_strand=True, circular=True),
          dict(mode=2, circular=True, gap_extend=-3, jump_score_same_contig_and_strand=-6, jump_score_inter_contig=-12)][case]
    alphabet = b"ACGT" if case % 2 == 0 else b"ACG"
    contigs = [gen.rand_seq(rng, l, alphabet) for l in (1500, 2305, 2560, 300, 769, 1024)]
    reads = [gen.chimeric_read(rng, contigs, rng.randint(300, 420), rng.randint(2, 4), strands=bool(kw.get("double_strand")),
                               wrap=True, alphabet=alphabet) for _ in range(3)]
    before = counters(emul_lib, 8)
    got, exp = run_both(oracle, emul_lib, kw, contigs, reads, 8, raw=False)
    compare(got, exp, f"special tiles case {case}")
    after = counters(emul_lib, 8)
    assert after[1] > before[1]
