"""CPU tier: the packed-key arithmetic of the fast fill (stitch_b200/csrc/dp_packed.h), run through the
CPU emulator with the kernel's structure (warp chunks, halo strips, special first/last tiles, hand-over
to the wide tail, wide re-fill for the traceback), must reproduce the oracle bit for bit."""
import random

import pytest

import gen
from stitch_b200._abi import make_opts
from test_emul_parity import compare, emul_lib, run_both  # noqa: F401


@pytest.mark.parametrize("strip", [1, 2, 8])
@pytest.mark.parametrize("block", range(6))
def test_fuzz_packed_small(oracle, emul_lib, strip, block):
    for seed in range(block * 50, block * 50 + 50):
        alphabet = [b"ACGT", b"AC", b"A", b"ACGTN"][seed % 4]
        contigs, reads = gen.fuzz_case(seed + 7000, max_contigs=5, max_len=90, max_read=70, alphabet=alphabet)
        kw = gen.fuzz_opts_packed(seed, strip)
        got, exp = run_both(oracle, emul_lib, kw, contigs, reads, strip, raw=(seed % 3 != 0))
        compare(got, exp, f"seed {seed} strip {strip} {kw}")


@pytest.mark.parametrize("strip", [1, 2, 8])
@pytest.mark.parametrize("block", range(4))
def test_fuzz_packed_multi_tile(oracle, emul_lib, strip, block):
    """Contigs spanning several warp tiles and warp chunks (halo strips), reads long enough for several
    checkpoint blocks and for the column base to drift."""
    tile = 32 * strip
    for seed in range(block * 8, block * 8 + 8):
        rng = random.Random(9000 + seed * 31 + strip)
        alphabet = [b"ACGT", b"AC", b"ACG"][seed % 3]
        lens = [rng.randint(1, 5 * tile) for _ in range(rng.randint(1, 5))] + [tile + 1, tile, tile - 1][: seed % 4]
        contigs = [gen.rand_seq(rng, l, alphabet) for l in lens]
        reads = [gen.chimeric_read(rng, contigs, rng.randint(20, 260), rng.randint(1, 4), strands=True,
                                   wrap=rng.random() < 0.5, noise=rng.random() < 0.8, alphabet=alphabet) for _ in range(3)]
        reads = [r if r else b"A" for r in reads]
        kw = gen.fuzz_opts_packed(seed + 100 * block, strip)
        got, exp = run_both(oracle, emul_lib, kw, contigs, reads, strip, raw=(seed % 2 == 0))
        compare(got, exp, f"seed {seed} strip {strip} {kw}")


def test_packed_plan_regime():
    """pk_plan must accept the reference CLI defaults at the benchmark shapes (checked in C by the emulator
    library through the parity above); here: the scorings gen.fuzz_opts_packed draws are inside the regime."""
    for strip in (1, 2, 8):
        for seed in range(200):
            kw = gen.fuzz_opts_packed(seed, strip)
            o = make_opts(**kw)
            sub = (o.match_score, o.mismatch_score)
            band = max(max(sub), 0) - min(sub) - min(o.jump_same, o.jump_opp, o.jump_inter)
            assert o.gap_extend < 0 and band // -o.gap_extend + 1 <= strip
