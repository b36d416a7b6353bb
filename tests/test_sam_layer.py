"""CPU tier: the product's SAM record layer (stitch_b200/csrc/host_sam.hpp, through the C ABI entry stitch_format_sam as
exported by the CPU emulator library) against the pure-Python restatement of the reference's SubAlignmentBuilder /
SamRecordFormatter (oracle/sam_oracle.py), on the chains the restated aligner (oracle/) produces for the same reads."""
import os
import random
import sys

import pytest

import gen
from stitch_b200._abi import make_opts
from test_emul_parity import emul_lib  # noqa: F401

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import sam_oracle  # noqa: E402


def oracle_sam(oracle, kw, named, reads, headers, quals, sam_opts):
    o = make_opts(**kw)
    chains, _ = oracle.OracleAligners(o, named).batch(reads, raw=False)
    targets = [(n, len(s)) for n, s in named]
    scoring = (o.match_score, o.mismatch_score, o.gap_open, o.gap_extend)
    return [sam_oracle.format_sam(headers[r], bytes(reads[r]), None if quals is None else quals[r], chains[r], targets, scoring,
                                  **sam_opts) for r in range(len(reads))]


@pytest.mark.parametrize("block", range(4))
def test_sam_records_match_restatement(oracle, emul_lib, block):
    for seed in range(block * 30, block * 30 + 30):
        rng = random.Random(5000 + seed)
        alphabet = [b"ACGT", b"AC", b"ACGTN"][seed % 3]
        contigs, reads = gen.fuzz_case(seed + 400, max_contigs=4, max_len=90, max_read=90, alphabet=alphabet)
        kw = gen.fuzz_opts_packed(seed, 8) if seed % 2 else gen.fuzz_opts(seed)
        kw["suboptimal"] = rng.random() < 0.5
        kw["suboptimal_pct"] = rng.choice([0.0, 20.0, 90.0])
        named = [(f"ctg{k}", s) for k, s in enumerate(contigs)]
        headers = [f"read{r}/x some comment {r}" for r in range(len(reads))]
        quals = None if seed % 4 == 0 else [bytes(rng.randrange(33, 74) for _ in r) for r in reads]
        sam_opts = dict(soft_clip=rng.random() < 0.5, use_eq_and_x=rng.random() < 0.5, pick_primary=rng.randrange(2),
                        filter_secondary=rng.random() < 0.4, filter_secondary_pct=rng.choice([0.0, 10.0, 50.0, 100.0]))
        e = emul_lib.EmulAligners(make_opts(**kw), named, strip=8)
        try:
            _, got = e.batch_sam(reads, headers, quals, sam_opts)
        except Exception as ex:   # a chain without operations: the reference panics there
            assert "without operations" in str(ex)
            with pytest.raises(IndexError):
                oracle_sam(oracle, kw, named, reads, headers, quals, sam_opts)
            continue
        exp = oracle_sam(oracle, kw, named, reads, headers, quals, sam_opts)
        assert got == exp, f"seed {seed} {kw} {sam_opts}"


def test_sam_record_example(oracle, emul_lib):
    """A hand-checked record: 25 matches on the forward strand of the only contig (the reference's API test case shape)."""
    contig = b"ACGTTGCATGCAAGTCCGATTAGCAGGCTTAACG"
    read = contig[4:29]
    e = emul_lib.EmulAligners(make_opts(), [("chr1", contig)], strip=8)
    chains, sam = e.batch_sam([read], ["r1 desc"], [b"I" * len(read)], None)
    f = sam[0][0].split("\t")
    assert f[:9] == ["r1", "0", "chr1", "5", "60", "25M", "*", "0", "0"]
    assert f[9] == read.decode() and f[10] == "I" * 25
    assert f[11:] == ["qs:i:0", "qe:i:25", "ts:i:4", "te:i:29", "as:i:25", "si:i:0", "sc:Z:25M", "cl:i:1", "ci:i:0", "cn:i:1",
                      "AS:i:25", "NM:i:0", "SA:Z:chr1,5,+,25M,60,0"]
