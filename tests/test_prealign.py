"""Pre-alignment contig selection (stitch_b200/csrc/prealign_core.h) and contig tables beyond the
reference's 256 contig-strands.

The reference delegates the pre-alignment to the `bio` crate (banded Smith-Waterman seeded by a k-mer hash,
mod.rs:556-604), which is not part of the reference checkout and is exercised by no reference test: parity unpinned.  What
is tested here is the contract around it, which IS the reference's (mod.rs:243-295, 338-339):
  * with `-x` a read is aligned to the selected contig-strands only, and its chains equal the oracle's chains for the same
    subset, bit for bit;
  * every contig-strand a read's segments were drawn from is selected (recall on synthetic reads);
  * a read that reaches the minimum score nowhere is not aligned at all: no chain, no score, an unmapped record;
  * without `-x` every contig is aligned and the score is the one of the first passing target;
  * the Option<i32> score comes back through stitch_results_prealign and lands in the `xs` tag.
CPU tier: the product's host driver over the emulator backend (which implements prealign_core.h sequentially); the gpu
tier (test_gpu_parity.py) repeats it on the CUDA kernel and compares the selections with the emulator's."""
import random

import pytest

import gen
from stitch_b200._abi import make_opts
from test_emul_parity import emul_lib  # noqa: F401


def make_case(seed, n_contigs, lo, hi, n_reads, read_len, nseg, flips=True):
    rng = random.Random(seed)
    contigs = [gen.rand_seq(rng, rng.randint(lo, hi)) for _ in range(n_contigs)]
    reads, truth = [], []
    for _ in range(n_reads):
        parts, t = [], set()
        for _s in range(nseg):
            c = rng.randrange(n_contigs)
            l = min(read_len // nseg, len(contigs[c]))
            start = rng.randrange(0, len(contigs[c]) - l + 1)
            piece = contigs[c][start:start + l]
            flip = flips and rng.random() < 0.5
            parts.append(gen.revcomp(piece) if flip else piece)
            t.add(c + (n_contigs if flip else 0))
        reads.append(gen.noisy(rng, b"".join(parts)))
        truth.append(sorted(t))
    return contigs, reads, truth


def reduced_oracle(oracle, kw, contigs, read, strands):
    """The oracle holds at most 256 contig-strands (8-bit index, like the reference): build it over the forward contigs the
    selected strands belong to (same relative order, so the same tie-breaks) and map the indices back."""
    T = len(contigs)
    involved = sorted({s % T for s in strands})
    rank = {g: k for k, g in enumerate(involved)}
    Tr = len(involved)
    fwd = lambda s: rank[s % T] + (Tr if s >= T else 0)
    back = {fwd(s): s for s in strands}
    okw = {k: v for k, v in kw.items() if not k.startswith("pre_align") and k not in ("kmer_size", "band_width")}
    exp, _ = oracle.OracleAligners(make_opts(**okw), [(f"c{g}", contigs[g]) for g in involved]).batch(
        [read], subsets=[sorted(back)], raw=False)
    for a in exp[0]:
        a.start_contig_idx = back[a.start_contig_idx]
        a.end_contig_idx = back[a.end_contig_idx]
        a.ops = [(k, back[x] if k == 6 else x, y) for k, x, y in a.ops]
    return exp[0]


def check_subset_mode(oracle, make_aligner, kw, contigs, reads, truth):
    """`make_aligner(kw, named)` -> object with batch(reads), prealign_batch(reads), last_prealign_scores, close()."""
    named = [(f"c{k}", s) for k, s in enumerate(contigs)]
    al = make_aligner(kw, named)
    got = al.batch(reads)
    pre = al.last_prealign_scores
    sel, best = al.prealign_batch(reads)
    al.close()
    for r in range(len(reads)):
        assert set(truth[r]) <= set(sel[r]), (r, truth[r], sel[r])            # recall
        assert 0 < len(sel[r]) <= 256 and sel[r] == sorted(sel[r])
        assert pre[r] == best[r] and best[r] >= kw.get("pre_align_min_score", 100)
        exp = reduced_oracle(oracle, kw, contigs, reads[r], sel[r])
        assert [a.key() for a in got[r]] == [a.key() for a in exp], r
    return sel, best


def emul_maker(emul_lib):
    return lambda kw, named: emul_lib.EmulAligners(make_opts(**kw), named, strip=8)


def test_prealign_subset_equals_oracle_on_the_same_subset(oracle, emul_lib):
    kw = dict(double_strand=True, pre_align=True, kmer_size=8, band_width=20, pre_align_min_score=40)
    contigs, reads, truth = make_case(11, 12, 300, 600, 6, 400, 3)
    check_subset_mode(oracle, emul_maker(emul_lib), kw, contigs, reads, truth)
    # circular + suboptimal chains, single strand, default k = 12
    kw = dict(circular=True, suboptimal=True, pre_align=True, band_width=50, pre_align_min_score=60)
    contigs, reads, truth = make_case(12, 8, 300, 600, 4, 500, 2, flips=False)
    check_subset_mode(oracle, emul_maker(emul_lib), kw, contigs, reads, truth)


def test_table_beyond_256_contig_strands(oracle, emul_lib):
    """300 contigs x 2 strands = 600 contig-strands in the table (the reference asserts <= 256,
    packed_length_cell.rs:139-140); the pre-alignment keeps every read below 256."""
    kw = dict(double_strand=True, pre_align=True, kmer_size=10, band_width=30, pre_align_min_score=50)
    contigs, reads, truth = make_case(13, 300, 120, 260, 5, 360, 3)
    sel, _ = check_subset_mode(oracle, emul_maker(emul_lib), kw, contigs, reads, truth)
    assert any(max(s) >= 256 for s in sel)
    # explicit subsets work on the big table too; no subset at all is refused (more than 256 contig-strands for one read)
    named = [(f"c{k}", s) for k, s in enumerate(contigs)]
    e = emul_lib.EmulAligners(make_opts(double_strand=True), named, strip=8)
    got = e.batch(reads[:2], subsets=sel[:2])
    for r in range(2):
        assert [a.key() for a in got[r]] == [a.key() for a in reduced_oracle(oracle, kw, contigs, reads[r], sel[r])]
    with pytest.raises(emul_lib.EmulError, match="more than 256 contig-strands"):
        e.batch(reads[:1])
    e.close()


def test_reads_that_reach_the_score_nowhere_are_not_aligned(emul_lib):
    rng = random.Random(5)
    contigs = [gen.rand_seq(rng, 400) for _ in range(4)]
    named = [(f"c{k}", s) for k, s in enumerate(contigs)]
    good = contigs[2][50:350]
    junk = gen.rand_seq(rng, 300)
    short = contigs[1][10:60]                                   # 50 matching bases < min score 100
    e = emul_lib.EmulAligners(make_opts(double_strand=True, pre_align=True), named, strip=8)
    chains, sam = e.batch_sam([good, junk, short, good.lower()], ["g", "j", "s", "gl"], None, None)
    assert [len(c) for c in chains] == [1, 0, 0, 1]
    assert e.last_prealign_scores[0] == 300 and e.last_prealign_scores[1] is None and e.last_prealign_scores[2] is None
    assert sam[1] == ["j\t4\t*\t0\t0\t*\t*\t0\t0\t" + junk.decode() + "\t*"]           # (Vec::new(), None): no xs tag, mod.rs:282-284
    assert "\txs:i:300\t" in sam[0][0] and sam[0][0].split("\t")[2] == "c2"             # the score is the suboptimal score, mod.rs:678-687
    assert chains[3][0].key() == chains[0][0].key()
    e.close()


def test_without_subset_contigs_every_contig_is_aligned(oracle, emul_lib):
    """`-p` with --pre-align-subset-contigs false: the pre-alignment only gates the read (mod.rs:274-277, 293-295); the score is
    the best strand of the FIRST passing target."""
    rng = random.Random(6)
    contigs = [gen.rand_seq(rng, 500) for _ in range(5)]
    named = [(f"c{k}", s) for k, s in enumerate(contigs)]
    read = contigs[3][100:300] + gen.revcomp(contigs[1][50:400])       # targets 1 (reverse, 350 b) and 3 (forward, 200 b) pass
    kw = dict(double_strand=True, pre_align=True, pre_align_subset_contigs=False)
    e = emul_lib.EmulAligners(make_opts(**kw), named, strip=8)
    got = e.batch([read, gen.rand_seq(rng, 200)])
    assert e.last_prealign_scores == [350, None] and got[1] == []
    exp, _ = oracle.OracleAligners(make_opts(double_strand=True), named).batch([read], raw=False)
    assert [a.key() for a in got[0]] == [a.key() for a in exp[0]]
    e.close()
