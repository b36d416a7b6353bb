#!/usr/bin/env python3
"""Extracts the reference's known-answer tests for the alignment path into JSON fixtures.

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/extract_reference_kats.py

It PARSES the reference's own Rust test modules (no reference code is copied into the repo, only
the input/expected-output vectors of its tests):

  * fg-stitch-lib/src/align/aligners/single_contig_aligner.rs:877-1774  -> sca_kats.json (63 cases)
  * fg-stitch-lib/src/align/aligners/multi_contig_aligner.rs:392-738    -> mca_kats.json (9 tests, 13 assertions)
  * fg-stitch-lib/src/align/alignment.rs:395-410,586-707                -> split_at_y_kats.json (7 cases)
  * fg-stitch-lib/src/align/aligners/mod.rs:984-1003                     -> api_kats.json (1 case)

The multi-contig, split_at_y and API cases are few and irregular, so they are transcribed by hand
below with their line numbers and cross-checked against the source text by `_check_present`.
"""
import json
import os
import re
import sys

REF = os.environ.get("STITCH_REFERENCE", "/root/reference")
LIB = os.path.join(REF, "fg-stitch-lib", "src")
OUT = os.path.dirname(os.path.abspath(__file__))


def clean(seq: str) -> str:
    """The tests' `s()` helper: strip '-', ' ', '_' and upper-case."""
    return "".join(ch for ch in seq if ch not in "- _").upper()


def rust_int(expr: str) -> int:
    expr = expr.replace("_", "").strip()
    if not re.fullmatch(r"[0-9+\-*() \n]+", expr):
        raise ValueError(f"unexpected expression: {expr!r}")
    return int(eval(expr))  # arithmetic on integer literals only (checked above)


def split_args(s: str):
    """Splits a Rust argument list on top-level commas."""
    args, depth, cur, in_str = [], 0, "", False
    for ch in s:
        if ch == '"':
            in_str = not in_str
        if not in_str:
            if ch in "([":
                depth += 1
            elif ch in ")]":
                depth -= 1
            elif ch == "," and depth == 0:
                args.append(cur.strip())
                cur = ""
                continue
        cur += ch
    if cur.strip():
        args.append(cur.strip())
    return args


def extract_sca():
    path = os.path.join(LIB, "align", "aligners", "single_contig_aligner.rs")
    text = open(path).read()
    lines = text.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("pub mod tests"))
    body = "\n".join(lines[start:])
    offset = start  # 0-based line index of the tests module
    cases = []
    for m in re.finditer(r"#\[rstest\]\s*\n\s*fn (test_\w+)\(\) \{(.*?)\n    \}", body, re.S):
        name, blk = m.group(1), m.group(2)
        line = offset + body[: m.start(1)].count("\n") + 1
        xs = re.search(r'let x = s\("([^"]*)"\)', blk)
        ys = re.search(r'let y = s\("([^"]*)"\)', blk)
        assert xs and ys, name
        match, mismatch = 1, -1
        gap_open, gap_extend, jump = -5, -1, -10  # SingleContigAligner::default(), SCA:85-90
        mp = re.search(r"MatchParams::new\(([^,]+),([^)]+)\)", blk)
        if mp:
            match, mismatch = rust_int(mp.group(1)), rust_int(mp.group(2))
        nw = re.search(r"SingleContigAligner::new\(([^;]*?)match_fn\)", blk, re.S)
        if nw:
            a = split_args(nw.group(1))
            gap_open, gap_extend, jump = rust_int(a[0]), rust_int(a[1]), rust_int(a[2])
        else:
            assert "SingleContigAligner::default()" in blk, name
        sj = re.search(r"set_jump_score\(([^)]+)\)", blk)
        if sj:
            jump = rust_int(sj.group(1))
        go = re.search(r"scoring\.gap_open = ([^;]+);", blk)
        if go:
            gap_open = rust_int(go.group(1))
        circular = "set_circular(true)" in blk
        md = re.search(r"aligner\.(global|querylocal|targetlocal|local)\(&x, &y\)", blk)
        assert md, name
        aa = re.search(r"assert_alignment\((.*?)\);", blk, re.S)
        args = split_args(aa.group(1))
        assert args[0] == "&alignment" and len(args) == 8, (name, args)
        cases.append({
            "name": name, "ref_line": line,
            "mode": {"local": 0, "querylocal": 1, "targetlocal": 2, "global": 3}[md.group(1)],
            "x": clean(xs.group(1)), "y": clean(ys.group(1)),
            "match": match, "mismatch": mismatch, "gap_open": gap_open, "gap_extend": gap_extend,
            "jump": jump, "circular": circular,
            "expect": {
                "xstart": rust_int(args[1]), "xend": rust_int(args[2]),
                "ystart": rust_int(args[3]), "yend": rust_int(args[4]),
                "score": rust_int(args[5]), "start_contig_idx": 0,
                "cigar": json.loads(args[6]), "length": rust_int(args[7]),
            },
        })
    return cases


MIN = -858993459


def rc(s: str) -> str:
    comp = dict(zip("AGCTYRWSKMDVHBN", "TCGARYWSMKHBDVN"))
    return "".join(comp.get(c, c) for c in reversed(s))


def mca_cases():
    """multi_contig_aligner.rs tests, transcribed (line = `fn test_*`)."""
    def G(x, o, e, j):  # scoring_global_custom :437-447
        return dict(match=1, mismatch=x, gap_open=o, gap_extend=e, jump_same=j, jump_opp=j, jump_inter=j,
                    xclip_prefix=MIN, xclip_suffix=MIN, yclip_prefix=MIN, yclip_suffix=MIN)

    def L(x, o, e, j):  # scoring_local_custom :453-463
        d = G(x, o, e, j)
        d.update(xclip_prefix=0, xclip_suffix=0, yclip_prefix=0, yclip_suffix=0)
        return d

    def two(x):
        return [dict(name="fwd", fwd=True, seq=x), dict(name="revcomp", fwd=False, seq=rc(x))]

    def exp(xs, xe, ys, ye, score, c, cigar, length):
        return dict(xstart=xs, xend=xe, ystart=ys, yend=ye, score=score, start_contig_idx=c, cigar=cigar, length=length)

    E5 = -100000
    cases = [
        dict(name="test_identical", ref_line=467, contigs=two("ACGTAACC"), scoring=G(-1, -5, -1, -10),
             y="ACGTAACC", expect=exp(0, 8, 0, 8, 8, 0, "8=", 8)),
        dict(name="test_identical_revcomp", ref_line=480, contigs=two("ACGTAACC"), scoring=G(-1, -5, -1, -10),
             y=rc("ACGTAACC"), expect=exp(0, 8, 0, 8, 8, 1, "8=", 8)),
        dict(name="test_fwd_to_fwd_jump", ref_line=492, contigs=two("AAGGCCTT"), scoring=G(-1, E5, E5, -1),
             y="AACCGGTT", expect=exp(0, 8, 0, 8, 5, 0, "2=2J2=4j2=2J2=", 8)),
        dict(name="test_fwd_to_rev_jump", ref_line=526, contigs=two("AACCTTGG"), scoring=G(E5, E5, E5, -1),
             y="AACCGGTT", expect=exp(0, 8, 0, 8, 7, 0, "4=1C0J4=", 8)),
        dict(name="test_rev_to_fwd_jump", ref_line=550, contigs=two("CCAAGGTT"), scoring=G(E5, E5, E5, -1),
             y="AACCGGTT", expect=exp(0, 8, 0, 8, 7, 1, "4=1c0J4=", 8)),
        dict(name="test_fwd_to_rev_long_jump", ref_line=574, contigs=two("AACCAAAATTGG"), scoring=G(E5, E5, E5, -1),
             y="AACCGGTT", expect=exp(0, 12, 0, 8, 7, 0, "4=1C4J4=", 8)),
        dict(name="test_rev_to_fwd_long_jump", ref_line=603, contigs=two("CCAANNNNGGTT"), scoring=G(E5, E5, E5, -1),
             y="AACCGGTT", expect=exp(0, 12, 0, 8, 7, 1, "4=1c4J4=", 8)),
        dict(name="test_many_contigs", ref_line=627,
             contigs=[dict(name=f"contig-{i}", fwd=True, seq=s) for i, s in enumerate(
                 ["TATATCCCCCTATATATATATATATATA", "ATATATTATATATATATATATATGGGGG", "AAAAA", "TTTTTTTTTTTTTTTT"])],
             scoring=L(E5, E5, E5, -1), y="AAAAACCCCCGGGGGAAAAATTTTTTTTTTTTTTTT",
             expect=exp(0, 16, 0, 36, 32, 2, "5=2c0J5=1C13J5=1C28j5=1C5j16=", 36)),
    ]
    x1 = "AAAAATTTTTAAAAA"
    js_contigs = [dict(name="chr1", fwd=True, seq=x1), dict(name="chr1", fwd=False, seq=rc(x1)),
                  dict(name="chr2", fwd=True, seq="AAAAA")]
    for k, (js, e) in enumerate([
        ((-1, -2, -2), exp(0, 15, 0, 10, 9, 0, "5=5J5=", 10)),
        ((-2, -1, -2), exp(5, 15, 0, 10, 9, 1, "5A5=1c5j5=", 10)),
        ((-2, -2, -1), exp(0, 15, 0, 10, 9, 2, "5=2c5J5=", 10)),
        ((-1, -1, -1), exp(0, 15, 0, 10, 9, 0, "5=5J5=", 10)),
        ((-2, -1, -1), exp(5, 15, 0, 10, 9, 1, "5A5=1c5j5=", 10)),
    ]):
        sc = L(-1, E5, E5, -1)
        sc.update(jump_same=js[0], jump_opp=js[1], jump_inter=js[2])
        cases.append(dict(name=f"test_jump_scores[{k}]", ref_line=669, contigs=js_contigs, scoring=sc,
                          y="AAAAAAAAAA", expect=e))
    return cases


def split_cases():
    """alignment.rs:586-707 fixtures and the `test_split_at_y` table (:680-686)."""
    M, XJ, YC = 0, 6, 5
    def aln(xs, xe, xl, ys, ye, yl, ops, mode):
        return dict(score=0, xstart=xs, xend=xe, xlen=xl, ystart=ys, yend=ye, ylen=yl, start_contig_idx=0,
                    end_contig_idx=0, ops=ops, mode=mode, length=10)
    m5 = [[M, 5, 0]]
    no_y_jump = lambda: aln(45, 5, 50, 0, 10, 10, m5 + [[XJ, 0, 0]] + m5, 0)
    slop5 = lambda: aln(40, 10, 50, 0, 10, 10, m5 + [[XJ, 0, 5]] + m5, 0)
    clip = lambda mode: aln(40, 10, 50, 0, 20, 20, m5 + [[YC, 5, 0], [XJ, 0, 5], [YC, 5, 0]] + m5, mode)
    empty = dict(score=0, xstart=0, xend=0, xlen=0, ystart=0, yend=0, ylen=0, start_contig_idx=0, end_contig_idx=0,
                 ops=[], mode=3, length=0)
    def exp(xs, xe, ys, ye, cigar, length):
        return dict(xstart=xs, xend=xe, ystart=ys, yend=ye, score=0, start_contig_idx=0, cigar=cigar, length=length)
    # modes: 0 Local, 1 QueryLocal, 2 TargetLocal, 3 Global
    return [
        dict(name="empty", ref_line=680, alignment=empty, y_pivot=0, expect=exp(0, 0, 0, 0, "", 0)),
        dict(name="no_y_jump", ref_line=681, alignment=no_y_jump(), y_pivot=5, expect=exp(0, 50, 0, 10, "5=40J5=", 10)),
        dict(name="slop_5_on_x", ref_line=682, alignment=slop5(), y_pivot=5, expect=exp(5, 45, 0, 10, "5=30J5=", 10)),
        dict(name="y_clipping_global", ref_line=683, alignment=clip(3), y_pivot=5, expect=exp(0, 50, 0, 20, "5A10B5=30J5=5A", 10)),
        dict(name="y_clipping_local", ref_line=684, alignment=clip(0), y_pivot=5, expect=exp(5, 45, 10, 20, "5=30J5=", 10)),
        dict(name="y_clipping_targetlocal", ref_line=685, alignment=clip(2), y_pivot=5, expect=exp(5, 45, 0, 20, "10B5=30J5=", 10)),
        dict(name="y_clipping_querylocal", ref_line=686, alignment=clip(1), y_pivot=5, expect=exp(0, 50, 10, 20, "5A5=30J5=5A", 10)),
    ]


def api_cases():
    seq = "ACGGACAGATCGAATACGACAGGAC"
    return [dict(name="test_case_insensitive", ref_line=985, contigs=[dict(name="test-contig", seq=seq)],
                 read=seq, expect=dict(n_chains=1, length=25, cigar="25="))]


def _check_present(relpath, needles):
    text = open(os.path.join(LIB, relpath)).read()
    for n in needles:
        if n not in text:
            raise SystemExit(f"{relpath}: expected to find {n!r} (reference changed?)")


def main():
    if not os.path.isdir(LIB):
        raise SystemExit(f"reference not found at {REF}")
    sca = extract_sca()
    assert len(sca) == 63, len(sca)
    _check_present("align/aligners/multi_contig_aligner.rs", [
        '"2=2J2=4j2=2J2="', '"4=1C0J4="', '"4=1c0J4="', '"4=1C4J4="', '"4=1c4J4="',
        '"5=2c0J5=1C13J5=1C28j5=1C5j16="', '"5=5J5="', '"5A5=1c5j5="', '"5=2c5J5="',
        's("CCAANNNNGGTT")', 's("AAAAATTTTTAAAAA")', "set_jump_scores(-2, -1, -1)"])
    _check_present("align/alignment.rs", ['"5A10B5=30J5=5A"', '"10B5=30J5="', '"5A5=30J5=5A"', '"5=40J5="', '"5=30J5="'])
    _check_present("align/aligners/mod.rs", ['b"ACGGACAGATCGAATACGACAGGAC"'])
    out = {"sca_kats.json": sca, "mca_kats.json": mca_cases(), "split_at_y_kats.json": split_cases(),
           "api_kats.json": api_cases()}
    for fn, data in out.items():
        with open(os.path.join(OUT, fn), "w") as f:
            json.dump(data, f, indent=1)
            f.write("\n")
        print(f"{fn}: {len(data)} cases")


if __name__ == "__main__":
    sys.exit(main())
