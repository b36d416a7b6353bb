"""Deterministic synthetic workloads of BASELINE.json's configs.

Contigs are iid uniform ACGT; a read is a concatenation of uniformly placed substrings of random
contigs (strand flipped with p = 0.5 and segments allowed to wrap the contig origin when
`strands` / `wrap` are set), followed by ONT-like noise: 3 % substitutions, 1 % insertions,
1 % deletions with geometric indel lengths (p = 0.7).  Seeds: 20240 + config number.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[_ACGT] = np.frombuffer(b"TGCA", dtype=np.uint8)


def _rand_bases(rng, n):
    return _ACGT[rng.integers(0, 4, size=n)]


def make_contigs(rng, count, lo, hi):
    return [_rand_bases(rng, int(rng.integers(lo, hi + 1))) for _ in range(count)]


def _noise(rng, s, sub=0.03, ins=0.01, dele=0.01, p_geo=0.7):
    out = []
    i = 0
    n = len(s)
    r = rng.random(n)
    while i < n:
        if r[i] < dele:
            i += int(rng.geometric(p_geo))
            continue
        if r[i] < dele + ins:
            out.append(_rand_bases(rng, int(rng.geometric(p_geo))))
        if rng.random() < sub:
            out.append(_rand_bases(rng, 1))
        else:
            out.append(s[i:i + 1])
        i += 1
    return np.concatenate(out) if out else _rand_bases(rng, 1)


def make_read(rng, contigs, length, seg_lo, seg_hi, strands=False, wrap=False, truth=None):
    """`truth` (a list) receives the contig-strand index of every segment: contig k forward = k, reverse = len(contigs) + k."""
    nseg = int(rng.integers(seg_lo, seg_hi + 1))
    cuts = np.sort(rng.integers(1, length, size=nseg - 1)) if nseg > 1 else np.array([], dtype=np.int64)
    bounds = np.concatenate(([0], cuts, [length]))
    parts = []
    for k in range(nseg):
        l = int(bounds[k + 1] - bounds[k])
        if l <= 0:
            continue
        ci = int(rng.integers(0, len(contigs)))
        c = contigs[ci]
        l = min(l, len(c))
        if wrap:
            start = int(rng.integers(0, len(c)))
            piece = np.concatenate((c, c))[start:start + l]
        else:
            start = int(rng.integers(0, len(c) - l + 1))
            piece = c[start:start + l]
        flipped = bool(strands and rng.random() < 0.5)
        if flipped:
            piece = _COMP[piece][::-1]
        if truth is not None:
            truth.append(ci + (len(contigs) if flipped else 0))
        parts.append(piece)
    clean = np.concatenate(parts)
    noisy = _noise(rng, clean)
    # keep the nominal read length (the configs quote reads "of 10 000 b")
    if len(noisy) >= length:
        return noisy[:length]
    return np.concatenate((noisy, _rand_bases(rng, length - len(noisy))))


def config(number, n_reads, read_len=None, truth=None):
    """Returns (opts kwargs, [(name, bytes)] contigs, [bytes] reads) for BASELINE config `number`.
    `truth` (a list) receives, per read, the contig-strand indices its segments were drawn from."""
    rng = np.random.default_rng(20240 + number)
    tr = [[] for _ in range(n_reads)]
    if number in (1, 2, 5):
        contigs = make_contigs(rng, 20, 7000, 9000)
        L = read_len or 10000
        two = number == 2
        reads = [make_read(rng, contigs, L, 3, 6, strands=two, wrap=two, truth=tr[k]) for k in range(n_reads)]
        kw = dict(double_strand=True, circular=True) if two else {}
    elif number == 6:   # not a BASELINE config: the floor of the quiet-tile skipping.  Real plasmid panels share backbones (origin,
        # resistance cassette): every contig = one common 4.8 kb backbone (60 %) with a unique 2.2-4.2 kb insert at the same
        # site, so that a read segment drawn from the backbone has 20 equally good contigs; reads as config 2
        backbone = _rand_bases(rng, 4800)
        contigs = [np.concatenate((backbone[:2400], _rand_bases(rng, int(rng.integers(2200, 4201))), backbone[2400:])) for _ in range(20)]
        L = read_len or 10000
        reads = [make_read(rng, contigs, L, 3, 6, strands=True, wrap=True, truth=tr[k]) for k in range(n_reads)]
        kw = dict(double_strand=True, circular=True)
    elif number == 3:   # 256-contig-strand slice of the construct database (the reference's own limit)
        contigs = make_contigs(rng, 128, 5000, 10000)
        reads = [make_read(rng, contigs, read_len or int(rng.integers(5000, 20001)), 3, 8, strands=True, truth=tr[k])
                 for k in range(n_reads)]
        kw = dict(double_strand=True)
    elif number == 4:
        contigs = make_contigs(rng, 50, 20000, 20000)
        reads = [make_read(rng, contigs, read_len or int(rng.integers(50000, 100001)), 10, 30, strands=True, truth=tr[k])
                 for k in range(n_reads)]
        kw = dict(double_strand=True)
    else:
        raise ValueError(number)
    if truth is not None:
        truth.extend(sorted(set(t)) for t in tr)
    named = [(f"contig{k}", c.tobytes()) for k, c in enumerate(contigs)]
    return kw, named, [r.tobytes() for r in reads]
