"""TEST INFRASTRUCTURE ONLY (the checker, never the product): pure-Python restatement of the reference's SAM record
layer, for small cases.

  SubAlignmentBuilder::build   fg-stitch-lib/src/align/sub_alignment.rs:170-241 (add_op :49-132, cmp_op :37-46,
                               swap_cigar :157-167)
  SamRecordFormatter::format   fg-stitch-lib/src/align/aligners/mod.rs:622-972 (header_to_name :612-619)
  reverse_complement           fg-stitch-lib/src/util/dna.rs:5-41

Works on the unit operations of a chain exactly as the reference does (one operation per base).  The byte-level
encoding of noodles 0.37 (BAM, its reading of FASTQ quality bytes) is not restated - parity unpinned there (SURVEY.md
section 8c): records are rendered as SAM text with the qualities passed through.  The reference has no tests for
this layer, so parity is by restatement only.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

MATCH, SUBST, DEL, INS, XCLIP, YCLIP, XJUMP, YJUMP = range(8)   # AlignmentOperation, constants.rs:20-29
MIN_SCORE = -858993459
_A, _B = b"AGCTYRWSKMDVHBN", b"TCGARYWSMKHBDVN"
_COMP = bytearray(range(256))
for _x, _y in zip(_A, _B):
    _COMP[_x] = _y
    _COMP[_x + 32] = _y + 32


def reverse_complement(s: bytes) -> bytes:
    return bytes(_COMP[c] for c in reversed(s))


@dataclass
class Sub:
    contig_idx: int = 0
    query_start: int = 0
    query_end: int = 0
    target_start: int = 0
    target_end: int = 0
    cigar: List[Tuple[str, int]] = field(default_factory=list)
    score: int = 0
    num_edits: int = 0


def cigar_str(c):
    return "".join(f"{n}{k}" for k, n in c)


def build_subs(chain, use_eq_and_x, scoring) -> List[Sub]:
    """chain: object with xstart, ystart, start_contig_idx and expanded_ops() -> [(kind, a, b)] unit operations;
    scoring: (match, mismatch, gap_open, gap_extend)."""
    match, mismatch, gap_open, gap_extend = scoring
    ops = chain.expanded_ops()
    st = dict(elements=[], q_start=chain.xstart, t_start=chain.ystart, q_off=chain.xstart, t_off=chain.ystart, score=0,
              num_edits=0, contig=chain.start_contig_idx)
    mk, xk = ("=", "X") if use_eq_and_x else ("M", "M")

    def cmp_op(last, cur):
        if use_eq_and_x:
            return last == cur
        return last == cur or (last[0] == SUBST and cur[0] == MATCH) or (last[0] == MATCH and cur[0] == SUBST)

    def snapshot():
        return Sub(st["contig"], st["q_start"], st["q_off"], st["t_start"], st["t_off"], list(st["elements"]), st["score"], st["num_edits"])

    def add_op(op, n):
        kind, a, b = op
        if kind == MATCH:
            st["score"] += match * n; st["q_off"] += n; st["t_off"] += n; st["elements"].append((mk, n)); return None
        if kind == SUBST:
            st["score"] += mismatch * n; st["q_off"] += n; st["t_off"] += n; st["elements"].append((xk, n)); return None
        if kind == DEL:
            st["score"] += gap_open + gap_extend * n; st["t_off"] += n; st["elements"].append(("D", n)); return None
        if kind == INS:
            st["score"] += gap_open + gap_extend * n; st["q_off"] += n; st["elements"].append(("I", n)); return None
        if kind == XJUMP:
            s = snapshot()
            st["elements"] = []; st["contig"] = a; st["t_start"] = st["t_off"]; st["q_start"] = b; st["q_off"] = b
            st["score"] = 0; st["num_edits"] = 0
            return s
        if kind == YJUMP:
            s = snapshot()
            st["elements"] = []; st["t_off"] += a; st["t_start"] = st["t_off"]; st["q_start"] = st["q_off"]
            st["score"] = 0; st["num_edits"] = 0
            return s
        assert n == 1
        return None

    out = []
    last, op_len = ops[0], 0          # IndexError on an empty chain, like the reference's panic
    for op in ops:
        if op[0] in (SUBST, INS, DEL):
            st["num_edits"] += 1
        if cmp_op(last, op):
            op_len += 1
        else:
            s = add_op(last, op_len)
            if s is not None and s.target_start < s.target_end:
                out.append(s)
            op_len = 1
        last = op
    s = add_op(last, op_len)
    out.append(s if s is not None else snapshot())
    swap = {"D": "I", "I": "D"}
    return [Sub(a.contig_idx, a.target_start, a.target_end, a.query_start, a.query_end,
                [(swap.get(k, k), n) for k, n in a.cigar], a.score, a.num_edits) for a in out]


def format_sam(header: str, bases: bytes, quals: Optional[bytes], chains, targets, scoring, pre_alignment_score=None,
               soft_clip=False, use_eq_and_x=False, pick_primary=0, filter_secondary=False, filter_secondary_pct=10.0) -> List[str]:
    """targets: [(name, length)] of the forward contigs.  Returns SAM text lines."""
    import struct
    f32 = lambda v: struct.unpack("f", struct.pack("f", v))[0]
    name = header.split()[0]
    T = len(targets)
    q = (lambda lo, hi, rev: "*") if quals is None else (lambda lo, hi, rev: ((quals[lo:hi][::-1] if rev else quals[lo:hi]).decode("latin1") or "*"))
    if not chains:
        line = "\t".join([name, "4", "*", "0", "0", "*", "*", "0", "0", bases.decode() or "*", q(0, len(bases), False)])
        if pre_alignment_score is not None:
            line += f"\txs:i:{pre_alignment_score}"
        return [line]
    cands = [c.score for c in chains[1:]]
    if pre_alignment_score is not None:
        cands.append(pre_alignment_score)
    suboptimal = max(cands) if cands else None
    records, primary_alignment_score = [], MIN_SCORE
    for chain_idx, chain in enumerate(chains):
        hard_clip = not soft_clip
        subs = build_subs(chain, use_eq_and_x, scoring)
        assert subs
        key = (lambda s: (s.query_end - s.query_start, s.score)) if pick_primary == 0 else (lambda s: (s.score, s.query_end - s.query_start))
        primary = 0
        for k, s in enumerate(subs):          # max_by_key: the last maximum
            if key(s) >= key(subs[primary]):
                primary = k
        if chain_idx == 0:
            primary_alignment_score = subs[primary].score
        if filter_secondary:
            min_score = f32(f32(f32(float(primary_alignment_score)) * f32(filter_secondary_pct)) / f32(100.0))
            kept = []
            for old, s in enumerate(subs):
                if old == primary:
                    primary = len(kept)
                if f32(float(s.score)) >= min_score:
                    kept.append(s)
            subs = kept
        lines, sa = [], []
        for sub_idx, s in enumerate(subs):
            is_supp, is_sec = sub_idx != primary, chain_idx > 0
            assert s.contig_idx < 2 * T
            fwd = s.contig_idx < T
            flags = (0 if fwd else 0x10) | (0x100 if is_sec else 0) | (0x800 if is_supp else 0)
            clip_seq = hard_clip and is_sec
            cig = list(s.cigar) if (fwd and not clip_seq) else list(reversed(s.cigar))
            lo, hi = (s.query_start, s.query_end) if clip_seq else (0, len(bases))
            seq = bases[lo:hi] if fwd else reverse_complement(bases[lo:hi])
            clip = "H" if clip_seq else "S"
            pre = s.query_start if fwd else len(bases) - s.query_end
            suf = len(bases) - s.query_end if fwd else s.query_start
            full = ([(clip, pre)] if pre > 0 else []) + cig + ([(clip, suf)] if suf > 0 else [])
            ref_id = s.contig_idx % T
            ref_start = s.target_start + 1 if fwd else targets[ref_id][1] - s.target_end + 1
            mapq = 60 if chain_idx == 0 else 0
            tags = [f"qs:i:{s.query_start}", f"qe:i:{s.query_end}", f"ts:i:{s.target_start}", f"te:i:{s.target_end}", f"as:i:{chain.score}"]
            if suboptimal is not None:
                tags.append(f"xs:i:{suboptimal}")
            tags += [f"si:i:{sub_idx}", f"sc:Z:{cigar_str(cig)}", f"cl:i:{len(subs)}", f"ci:i:{chain_idx}", f"cn:i:{len(chains)}",
                     f"AS:i:{s.score}", f"NM:i:{s.num_edits}"]
            lines.append("\t".join([name, str(flags), targets[ref_id][0], str(ref_start), str(mapq), cigar_str(full) or "*", "*", "0", "0",
                                    seq.decode() or "*", q(lo, hi, not fwd)] + tags))
            sa.append(f"{targets[ref_id][0]},{ref_start},{'+' if fwd else '-'},{cigar_str(full)},{mapq},{s.num_edits}")
        if sa:
            r = primary % len(sa)
            sa = sa[-r:] + sa[:-r] if r else sa          # rotate_right(primary)
        records += [l + "\tSA:Z:" + ";".join(sa) for l in lines]
    return records
