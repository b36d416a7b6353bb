"""The reference's 13 multi-contig known-answer assertions (multi_contig_aligner.rs:467-737, tests/golden/mca_kats.json)
through the product's C ABI: on the CPU emulator of the kernels (always) and on the CUDA path (gpu tier).

The reference tests build MultiContigAligner by hand: explicit (name, strand) contigs and raw clip scores.  Through the
ABI the same aligner is: the forward contigs + `double_strand` (strand order: all forward contigs, then their reverse
complements, mod.rs:186-205), mode global when every clip score is MIN_SCORE and local when every clip score is 0, and
`subset_words` to drop the strands the test does not add (test_jump_scores :669 has chr1+, chr1-, chr2+ only).  Contig
indices differ between the two orders, so the golden tuple is asserted after mapping the ABI's indices back to the
reference's (start contig and every Xjump target; the cigar string is then the reference's own).  In the two-strand
cases the reference names the strands differently ("fwd" / "revcomp"), which makes its jumps inter-contig where the ABI's
are opposite-strand: all three jump scores are equal in those cases."""
import json
import os

import pytest

import gen
from stitch_b200._abi import make_opts
from stitch_b200.alignment import Alignment
from test_emul_parity import emul_lib  # noqa: F401

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLD, "mca_kats.json")))
MIN_SCORE = -858993459


def through_abi(case):
    """-> (opts kwargs, [(name, seq)] forward contigs, subset of ABI strand indices, {ABI index: reference index})"""
    sc = case["scoring"]
    clips = {sc[k] for k in ("xclip_prefix", "xclip_suffix", "yclip_prefix", "yclip_suffix")}
    assert clips in ({0}, {MIN_SCORE})
    fw = [c for c in case["contigs"] if c["fwd"]]
    rv = [c for c in case["contigs"] if not c["fwd"]]
    if rv and len({sc["jump_same"], sc["jump_opp"], sc["jump_inter"]}) > 1:
        assert all(c["name"] in {f["name"] for f in fw} for c in rv)   # strands paired by name, as the ABI pairs them
    kw = dict(mode=0 if clips == {0} else 3, match_score=sc["match"], mismatch_score=sc["mismatch"], gap_open=sc["gap_open"],
              gap_extend=sc["gap_extend"], jump_score_same_contig_and_strand=sc["jump_same"],
              jump_score_same_contig_opposite_strand=sc["jump_opp"], jump_score_inter_contig=sc["jump_inter"], double_strand=bool(rv))
    abi2ref = {}
    for i, c in enumerate(case["contigs"]):
        if c["fwd"]:
            abi2ref[[f["name"] for f in fw].index(c["name"])] = i
        else:
            abi2ref[len(fw) + [gen.revcomp(f["seq"].encode()) for f in fw].index(c["seq"].encode())] = i
    return kw, [(c["name"], c["seq"].encode()) for c in fw], sorted(abi2ref), abi2ref


def assert_golden(a, case, abi2ref):
    e = case["expect"]
    in_ref = Alignment(a.score, a.xstart, a.xend, a.ystart, a.yend, a.xlen, a.ylen, abi2ref[a.start_contig_idx], abi2ref[a.end_contig_idx],
                       a.length, [(k, abi2ref[x] if k == 6 else x, y) for k, x, y in a.ops])
    got = (in_ref.xstart, in_ref.xend, in_ref.ystart, in_ref.yend, in_ref.score, in_ref.start_contig_idx, in_ref.cigar(), in_ref.length)
    assert got == (e["xstart"], e["xend"], e["ystart"], e["yend"], e["score"], e["start_contig_idx"], e["cigar"], e["length"]), case["name"]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['name']}@{c['ref_line']}")
def test_mca_kats_through_the_abi_on_the_emulator(emul_lib, case):
    kw, named, subset, abi2ref = through_abi(case)
    for strip in (8, 2):
        e = emul_lib.EmulAligners(make_opts(**kw), named, strip=strip)
        assert_golden(e.batch([case["y"].encode()], subsets=[subset], raw=True)[0][0], case, abi2ref)
        e.close()


@pytest.mark.gpu
def test_mca_kats_on_gpu():
    import stitch_b200
    for case in CASES:
        kw, named, subset, abi2ref = through_abi(case)
        for tuning in ({}, {"STITCH_PACKED": "0"}):   # the packed-key kernel where its regime allows, and the wide kernels
            os.environ.update(tuning)
            try:
                al = stitch_b200.Builder(**kw).build_aligners([stitch_b200.TargetSeq(n, s) for n, s in named], device=0)
            finally:
                for k in tuning:
                    os.environ.pop(k, None)
            assert_golden(al.custom_batch([case["y"].encode()], [subset])[0][0], case, abi2ref)
            al.close()
