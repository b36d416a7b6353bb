"""Deterministic synthetic inputs for the parity tests and the benchmark."""
import random

COMP = bytes.maketrans(b"ACGT", b"TGCA")


def revcomp(s: bytes) -> bytes:
    return s.translate(COMP)[::-1]


def rand_seq(rng, n, alphabet=b"ACGT"):
    return bytes(rng.choice(alphabet) for _ in range(n))


def noisy(rng, s: bytes, sub=0.03, ins=0.01, dele=0.01, alphabet=b"ACGT"):
    out = bytearray()
    for b in s:
        r = rng.random()
        if r < dele:
            while rng.random() < 0.3:
                pass
            continue
        if r < dele + ins:
            out.append(rng.choice(alphabet))
            while rng.random() < 0.3:
                out.append(rng.choice(alphabet))
        if rng.random() < sub:
            out.append(rng.choice(alphabet))
        else:
            out.append(b)
    return bytes(out)


def chimeric_read(rng, contigs, length, nseg, strands=False, wrap=False, noise=True, alphabet=b"ACGT"):
    """Concatenation of `nseg` substrings of random contigs (optionally flipped / origin-spanning)."""
    parts = []
    seg = max(1, length // nseg)
    for _ in range(nseg):
        c = rng.choice(contigs)
        l = min(seg, len(c))
        start = rng.randrange(0, len(c)) if wrap else rng.randrange(0, max(1, len(c) - l + 1))
        piece = (c + c)[start:start + l] if wrap else c[start:start + l]
        if strands and rng.random() < 0.5:
            piece = revcomp(piece)
        parts.append(piece)
    s = b"".join(parts)
    return noisy(rng, s, alphabet=alphabet) if noise else s


def fuzz_case(seed, max_contigs=5, max_len=60, max_read=60, alphabet=b"ACGT"):
    """Small adversarial case: low-complexity alphabets force score/length ties."""
    rng = random.Random(seed)
    nc = rng.randint(1, max_contigs)
    contigs = [rand_seq(rng, rng.randint(1, max_len), alphabet) for _ in range(nc)]
    reads = []
    for _ in range(rng.randint(1, 3)):
        if rng.random() < 0.7:
            r = chimeric_read(rng, contigs, rng.randint(1, max_read), rng.randint(1, 4), strands=rng.random() < 0.5,
                              wrap=rng.random() < 0.5, noise=rng.random() < 0.7, alphabet=alphabet)
        else:
            r = rand_seq(rng, rng.randint(1, max_read), alphabet)
        if not r:
            r = rand_seq(rng, 1, alphabet)
        reads.append(r)
    return contigs, reads


def fuzz_opts(seed):
    """Random scoring / mode / strand / circular options (as kwargs of _abi.make_opts)."""
    rng = random.Random(seed * 7919 + 13)
    kw = dict(mode=rng.randint(0, 3), double_strand=rng.random() < 0.5, circular=rng.random() < 0.5)
    style = rng.random()
    if style < 0.4:
        pass   # reference CLI defaults: A1 B-4 O-6 E-2 J-10
    elif style < 0.7:
        kw.update(match_score=rng.randint(1, 3), mismatch_score=-rng.randint(0, 5), gap_open=-rng.randint(0, 6),
                  gap_extend=-rng.randint(0, 3), default_jump_score=-rng.randint(0, 12))
    else:
        kw.update(match_score=1, mismatch_score=-rng.choice([1, 100000]), gap_open=-rng.choice([1, 5, 100000]),
                  gap_extend=-rng.choice([1, 100000]),
                  jump_score_same_contig_and_strand=-rng.randint(0, 3),
                  jump_score_same_contig_opposite_strand=-rng.randint(0, 3),
                  jump_score_inter_contig=-rng.randint(0, 3))
    kw["circular_slop"] = rng.choice([0, 2, 5, 20])
    return kw


def fuzz_opts_packed(seed, strip=8):
    """Scorings inside the packed kernel's regime (stitch_b200/csrc/dp_packed.h: insertion-chain reach
    <= strip rows, e < 0), all modes / strands / circular, with score ties made likely."""
    rng = random.Random(seed * 104729 + 7)
    kw = dict(mode=rng.randint(0, 3), double_strand=rng.random() < 0.5, circular=rng.random() < 0.5)
    if rng.random() < 0.3 and strip >= 8:
        pass   # reference CLI defaults
    else:
        match = rng.randint(0, 2)
        mismatch = -rng.randint(0, 4) if rng.random() < 0.9 else rng.randint(0, 1)
        jumps = [-rng.randint(0, 8) for _ in range(3)]
        band = max(match, mismatch, 0) - min(match, mismatch) - min(jumps)
        ext = band // strip + 1 + (rng.randint(0, 2) if rng.random() < 0.5 else 0)
        kw.update(match_score=match, mismatch_score=mismatch, gap_open=-rng.randint(0, 6), gap_extend=-ext,
                  jump_score_same_contig_and_strand=jumps[0], jump_score_same_contig_opposite_strand=jumps[1],
                  jump_score_inter_contig=jumps[2])
    kw["circular_slop"] = rng.choice([0, 2, 5, 20])
    return kw
