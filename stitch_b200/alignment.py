"""Host-side mirror of the reference's result type.

`Alignment` follows fg-stitch-lib/src/align/alignment.rs:16-51 (x = contig, y = read) and
`AlignmentOperation` follows aligners/constants.rs:20-85; ops are kept run-length encoded as
they cross the C ABI: (kind, a, b) with a = run length for Match/Subst/Del/Ins.
"""
from dataclasses import dataclass, field
from typing import List, Tuple

from ._abi import OP_DEL, OP_INS, OP_MATCH, OP_SUBST, OP_XCLIP, OP_XJUMP, OP_YCLIP, OP_YJUMP

Op = Tuple[int, int, int]


@dataclass
class Alignment:
    score: int = 0
    xstart: int = 0
    xend: int = 0
    ystart: int = 0
    yend: int = 0
    xlen: int = 0
    ylen: int = 0
    start_contig_idx: int = 0
    end_contig_idx: int = 0
    length: int = 0
    ops: List[Op] = field(default_factory=list)   # run-length encoded

    def key(self):
        """Everything the parity tests compare."""
        return (self.score, self.xstart, self.xend, self.ystart, self.yend, self.xlen, self.ylen,
                self.start_contig_idx, self.end_contig_idx, self.length, tuple(self.ops))

    def expanded_ops(self) -> List[Op]:
        out = []
        for kind, a, b in self.ops:
            if kind <= OP_INS:
                out.extend([(kind, 0, 0)] * a)
            else:
                out.append((kind, a, b))
        return out

    def cigar(self) -> str:
        """Alignment::cigar() (alignment.rs:105-149; letters: constants.rs:37-59)."""
        out = []
        contig, x = self.start_contig_idx, self.xstart
        letters = {OP_MATCH: "=", OP_SUBST: "X", OP_DEL: "D", OP_INS: "I"}
        # With run-length encoded input, consecutive equal non-special ops are already merged,
        # except across a Yjump (which is not "special" in the reference and breaks runs).
        pending_kind, pending_len = None, 0

        def flush():
            nonlocal pending_kind, pending_len
            if pending_len > 0:
                out.append(f"{pending_len}{letters[pending_kind]}")
            pending_kind, pending_len = None, 0

        for kind, a, b in self.ops:
            if kind <= OP_INS:
                if kind != pending_kind:
                    flush()
                    pending_kind = kind
                pending_len += a
                if kind != OP_DEL:
                    x += a
            elif kind == OP_YJUMP:
                # not special: it is an ordinary op with its own letter, run length counted per op
                flush()
                pending_kind = None
                out.append(f"1{a}S")
            else:
                flush()
                if kind == OP_XCLIP:
                    out.append(f"{a}A")
                    x += a
                elif kind == OP_YCLIP:
                    out.append(f"{a}B")
                else:  # Xjump(contig a, offset b)
                    s = ""
                    if a > contig:
                        s = f"{a - contig}C"
                    elif a < contig:
                        s = f"{contig - a}c"
                    s += f"{b - x}J" if b >= x else f"{x - b}j"
                    out.append(s)
                    x = b
                    contig = a
        flush()
        return "".join(out)

    def validate(self):
        """Alignment::validate() for mode Custom (alignment.rs:80-102)."""
        x, y, end, length = self.xstart, self.ystart, self.end_contig_idx, 0
        for kind, a, b in self.ops:
            if kind in (OP_MATCH, OP_SUBST):
                x += a; y += a; length += a
            elif kind == OP_INS:
                x += a; length += a
            elif kind == OP_DEL:
                y += a; length += a
            elif kind == OP_XCLIP:
                x += a
            elif kind in (OP_YCLIP, OP_YJUMP):
                y += a
            elif kind == OP_XJUMP:
                x = b
        assert y == self.yend, ("yend", y, self.yend)
        assert length == self.length, ("length", length, self.length)
        return True
