"""Pins the CPU oracle against every known-answer test the reference holds for the path
(SURVEY.md section 8c).  The fixtures are generated from the reference's own test modules by
tests/golden/extract_reference_kats.py."""
import json
import os

import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def check(a, e, cigar):
    assert a.xstart == e["xstart"], "xstart"
    assert a.xend == e["xend"], "xend"
    assert a.ystart == e["ystart"], "ystart"
    assert a.yend == e["yend"], "yend"
    assert a.score == e["score"], "score"
    assert a.start_contig_idx == e["start_contig_idx"], "start_contig_idx"
    assert cigar == e["cigar"], "cigar"
    assert a.length == e["length"], "length"


@pytest.mark.parametrize("case", load("sca_kats.json"), ids=lambda c: f"{c['name']}@{c['ref_line']}")
def test_single_contig_kats(oracle, case):
    a = oracle.sca(case["mode"], case["x"].encode(), case["y"].encode(), case["match"], case["mismatch"],
                   case["gap_open"], case["gap_extend"], case["jump"], case["circular"])
    check(a, case["expect"], a.oracle_cigar)
    # the Python mirror's cigar() agrees with the oracle's restatement of Alignment::cigar()
    assert a.cigar() == a.oracle_cigar


@pytest.mark.parametrize("case", load("mca_kats.json"), ids=lambda c: f"{c['name']}@{c['ref_line']}")
def test_multi_contig_kats(oracle, case):
    contigs = [dict(name=c["name"], fwd=c["fwd"], seq=c["seq"].encode()) for c in case["contigs"]]
    a = oracle.mca(contigs, case["scoring"], case["y"].encode())
    check(a, case["expect"], a.oracle_cigar)
    assert a.cigar() == a.oracle_cigar


@pytest.mark.parametrize("case", load("split_at_y_kats.json"), ids=lambda c: c["name"])
def test_split_at_y_kats(oracle, case):
    a = oracle.split_at_y(case["alignment"], case["y_pivot"])
    check(a, case["expect"], a.oracle_cigar)


@pytest.mark.parametrize("case", load("api_kats.json"), ids=lambda c: c["name"])
def test_api_kat(oracle, case):
    from stitch_b200._abi import make_opts
    al = oracle.OracleAligners(make_opts(), [(c["name"], c["seq"].encode()) for c in case["contigs"]])
    per_read, info = al.batch([case["read"].encode()])
    chains = per_read[0]
    assert len(chains) == case["expect"]["n_chains"]
    assert chains[0].length == case["expect"]["length"]
    assert chains[0].oracle_cigar == case["expect"]["cigar"]
    # lower-case input gives the same answer (test_case_insensitive, mod.rs:984-1003)
    per_read2, _ = al.batch([case["read"].lower().encode()])
    assert per_read2[0][0].key() == chains[0].key()
    assert info["cells"] == 25 * 25 and info["fills"] == 1
