"""CPU tier: the product's DP decomposition + host driver (run through the CPU emulator of the
kernel structure) must reproduce the oracle bit for bit: scores, coordinates, contigs, lengths
and every operation, including the reference's tie-breaking."""
import json
import os

import pytest

import gen
from stitch_b200._abi import make_opts

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def compare(per_read_a, per_read_b, ctx):
    assert len(per_read_a) == len(per_read_b)
    for r, (la, lb) in enumerate(zip(per_read_a, per_read_b)):
        assert len(la) == len(lb), f"{ctx} read {r}: chain count {len(la)} vs {len(lb)}"
        for k, (a, b) in enumerate(zip(la, lb)):
            assert a.key() == b.key(), f"{ctx} read {r} chain {k}:\n  emul   {a}\n  oracle {b}"


def run_both(oracle, emul_lib, kw, contigs, reads, strip, raw, subsets=None):
    named = [(f"c{k}", s) for k, s in enumerate(contigs)]
    opts = make_opts(**kw)
    o = oracle.OracleAligners(opts, named)
    e = emul_lib.EmulAligners(opts, named, strip=strip)
    exp, _ = o.batch(reads, subsets=subsets, raw=raw)
    got = e.batch(reads, subsets=subsets, raw=raw)
    return got, exp


@pytest.fixture(scope="module")
def emul_lib():
    import emul_lib as m
    m.build()
    return m


@pytest.mark.parametrize("strip", [1, 2, 8])
def test_sca_kats_through_emulator(oracle, emul_lib, strip):
    """The reference's single-contig KATs through the product's driver: custom_batch keeps the clip
    ops, so compare against the oracle's raw chain and check score/cigar-level fields of the KAT."""
    cases = json.load(open(os.path.join(GOLD, "sca_kats.json")))
    for c in cases:
        if c["gap_open"] != c["gap_open"]:
            continue
        kw = dict(mode=c["mode"], match_score=c["match"], mismatch_score=c["mismatch"], gap_open=c["gap_open"],
                  gap_extend=c["gap_extend"], default_jump_score=c["jump"], circular=c["circular"])
        got, exp = run_both(oracle, emul_lib, kw, [c["x"].encode()], [c["y"].encode()], strip, raw=True)
        compare(got, exp, c["name"])
        assert got[0][0].score == c["expect"]["score"], c["name"]
        assert got[0][0].length == c["expect"]["length"], c["name"]


@pytest.mark.parametrize("strip", [1, 2, 8])
@pytest.mark.parametrize("block", range(8))
def test_fuzz_custom(oracle, emul_lib, strip, block):
    """MultiContigAligner::custom_with_subset: random small cases, all modes/strands/circular."""
    for seed in range(block * 60, block * 60 + 60):
        alphabet = [b"ACGT", b"AC", b"A", b"ACGTN"][seed % 4]
        contigs, reads = gen.fuzz_case(seed, alphabet=alphabet)
        kw = gen.fuzz_opts(seed)
        got, exp = run_both(oracle, emul_lib, kw, contigs, reads, strip, raw=True)
        compare(got, exp, f"seed {seed} {kw}")


@pytest.mark.parametrize("strip", [2, 8])
@pytest.mark.parametrize("block", range(6))
def test_fuzz_align(oracle, emul_lib, strip, block):
    """Aligners::align: clip removal, circular origin re-alignment, sub-optimal chains."""
    import random
    for seed in range(1000 + block * 40, 1000 + block * 40 + 40):
        alphabet = [b"ACGT", b"AC", b"ACG"][seed % 3]
        contigs, reads = gen.fuzz_case(seed, max_contigs=4, max_len=80, max_read=80, alphabet=alphabet)
        kw = gen.fuzz_opts(seed)
        rng = random.Random(seed)
        kw["suboptimal"] = rng.random() < 0.5
        kw["suboptimal_pct"] = rng.choice([0.0, 20.0, 90.0])
        got, exp = run_both(oracle, emul_lib, kw, contigs, reads, strip, raw=False)
        compare(got, exp, f"seed {seed} {kw}")


@pytest.mark.parametrize("strip", [2, 8])
def test_fuzz_subsets(oracle, emul_lib, strip):
    import random
    for seed in range(3000, 3080):
        rng = random.Random(seed)
        contigs, reads = gen.fuzz_case(seed, max_contigs=5)
        kw = gen.fuzz_opts(seed)
        ns = len(contigs) * (2 if kw["double_strand"] else 1)
        subsets = []
        for _ in reads:
            sub = [c for c in range(ns) if rng.random() < 0.6]
            subsets.append(sub or [rng.randrange(ns)])
        got, exp = run_both(oracle, emul_lib, kw, contigs, reads, strip, raw=rng.random() < 0.5, subsets=subsets)
        compare(got, exp, f"seed {seed} {kw} {subsets}")


def test_medium_chimeric(oracle, emul_lib):
    """A few hundred bases, realistic noise, default scoring, both strands, circular."""
    import random
    rng = random.Random(42)
    contigs = [gen.rand_seq(rng, rng.randint(300, 700)) for _ in range(4)]
    reads = [gen.chimeric_read(rng, contigs, 400, 3, strands=True, wrap=True) for _ in range(4)]
    for kw in (dict(), dict(double_strand=True, circular=True), dict(mode=3), dict(mode=1, double_strand=True),
               dict(mode=2, circular=True, suboptimal=True)):
        got, exp = run_both(oracle, emul_lib, kw, contigs, reads, 8, raw=False)
        compare(got, exp, str(kw))
