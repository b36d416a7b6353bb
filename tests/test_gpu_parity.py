"""GPU tier: the CUDA path, called through the C ABI, must reproduce the oracle bit for bit
(scores, coordinates, contigs, lengths, every operation) on the same seeded inputs; plus
size-independent properties at larger sizes."""
import json
import os
import random

import pytest

import gen
from stitch_b200._abi import make_opts

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# Small checkpoint spacing / tracking window so that tiny fuzz reads still span several checkpoint
# blocks, re-filled units and window re-runs (the library reads these when a context is created).
TUNING = {"STITCH_CK_EVERY": "7", "STITCH_TRACK_WINDOW": "6"}
ALL_TUNING_KEYS = ("STITCH_CK_EVERY", "STITCH_TRACK_WINDOW", "STITCH_CLUSTER", "STITCH_CLUSTER_MIN_TILES", "STITCH_PACKED",
                   "STITCH_CLUSTER_SMEM", "STITCH_QUIET", "STITCH_WALK_IN_KERNEL")


def cluster_tuning(seed, base=None):
    """Thread-block clusters of 1 / 2 / 4 CTAs per read, also for layouts of a few tiles."""
    t = dict(base or {})
    t["STITCH_CLUSTER"] = str((1, 2, 4, 16, 8)[seed % 5])   # > 1: the clustered fill kernel + separate fix-up / walk kernels
    t["STITCH_CLUSTER_MIN_TILES"] = "1"
    t["STITCH_CLUSTER_SMEM"] = str(seed % 2)              # rolling state in the cluster's shared memory / in global memory
    if seed % 6 == 3:
        t["STITCH_WALK_IN_KERNEL"] = "0"                  # separate fix-up / walk kernels also for one CTA per read
    return t


def gpu_aligners(kw, named, max_inflight=0, tuning=None):
    import stitch_b200
    targets = [stitch_b200.TargetSeq(n, s) for n, s in named]
    for k in ALL_TUNING_KEYS:
        os.environ.pop(k, None)
    if tuning:
        os.environ.update(tuning)
    try:
        al = stitch_b200.Builder(**kw).build_aligners(targets, device=0)
    finally:
        for k in ALL_TUNING_KEYS:
            os.environ.pop(k, None)
    if max_inflight:
        al.set_max_inflight(max_inflight)
    return al


def compare(got, exp, ctx):
    assert len(got) == len(exp)
    for r, (la, lb) in enumerate(zip(got, exp)):
        assert len(la) == len(lb), f"{ctx} read {r}: chain count {len(la)} vs {len(lb)}"
        for k, (a, b) in enumerate(zip(la, lb)):
            assert a.key() == b.key(), f"{ctx} read {r} chain {k}:\n  gpu    {a}\n  oracle {b}"


def run_both(oracle, kw, contigs, reads, raw, subsets=None, max_inflight=0, tuning=None):
    named = [(f"c{k}", s) for k, s in enumerate(contigs)]
    exp, _ = oracle.OracleAligners(make_opts(**kw), named).batch(reads, subsets=subsets, raw=raw)
    al = gpu_aligners(kw, named, max_inflight, tuning)
    got = al.custom_batch(reads, subsets) if raw else al.align_batch(reads, subsets)
    al.close()
    return got, exp


def test_reference_kats_on_gpu(oracle):
    """The reference's own single- and multi-contig known-answer tests through the CUDA path."""
    for c in json.load(open(os.path.join(GOLD, "sca_kats.json"))):
        kw = dict(mode=c["mode"], match_score=c["match"], mismatch_score=c["mismatch"], gap_open=c["gap_open"],
                  gap_extend=c["gap_extend"], default_jump_score=c["jump"], circular=c["circular"])
        got, exp = run_both(oracle, kw, [c["x"].encode()], [c["y"].encode()], raw=True)
        compare(got, exp, c["name"])
        a = got[0][0]
        e = c["expect"]
        assert (a.score, a.length) == (e["score"], e["length"]), c["name"]
        if c["mode"] == 3:   # global: no clip ops are filtered by the mode wrapper, so the cigar is comparable
            assert (a.xstart, a.xend, a.ystart, a.yend, oracle.cigar_of(a)) == (e["xstart"], e["xend"], e["ystart"], e["yend"], e["cigar"])
    api = json.load(open(os.path.join(GOLD, "api_kats.json")))[0]
    al = gpu_aligners({}, [(c["name"], c["seq"].encode()) for c in api["contigs"]])
    chains = al.align_batch([api["read"].lower().encode()])[0]
    assert len(chains) == 1 and chains[0].length == 25 and chains[0].cigar() == "25="


@pytest.mark.parametrize("block", range(4))
def test_fuzz_custom_small(oracle, block):
    """Random small cases, all modes / strands / circular, low-complexity alphabets to force ties.
    Many reads per batch so the persistent-CTA queue is exercised."""
    for seed in range(block * 50, block * 50 + 50):
        alphabet = [b"ACGT", b"AC", b"A", b"ACGTN"][seed % 4]
        contigs, reads = gen.fuzz_case(seed, alphabet=alphabet)
        kw = gen.fuzz_opts(seed)
        got, exp = run_both(oracle, kw, contigs, reads, raw=True, tuning=TUNING if seed % 2 else None)
        compare(got, exp, f"seed {seed} {kw}")


@pytest.mark.parametrize("block", range(3))
def test_fuzz_align_small(oracle, block):
    for seed in range(1000 + block * 40, 1000 + block * 40 + 40):
        alphabet = [b"ACGT", b"AC", b"ACG"][seed % 3]
        contigs, reads = gen.fuzz_case(seed, max_contigs=4, max_len=80, max_read=80, alphabet=alphabet)
        kw = gen.fuzz_opts(seed)
        rng = random.Random(seed)
        kw["suboptimal"] = rng.random() < 0.5
        kw["suboptimal_pct"] = rng.choice([0.0, 20.0, 90.0])
        got, exp = run_both(oracle, kw, contigs, reads, raw=False, tuning=TUNING if seed % 2 else None)
        compare(got, exp, f"seed {seed} {kw}")


def test_fuzz_subsets(oracle):
    for seed in range(3000, 3040):
        rng = random.Random(seed)
        contigs, reads = gen.fuzz_case(seed, max_contigs=5)
        kw = gen.fuzz_opts(seed)
        ns = len(contigs) * (2 if kw["double_strand"] else 1)
        subsets = []
        for _ in reads:
            sub = [c for c in range(ns) if rng.random() < 0.6]
            subsets.append(sub or [rng.randrange(ns)])
        got, exp = run_both(oracle, kw, contigs, reads, raw=rng.random() < 0.5, subsets=subsets,
                            tuning=TUNING if seed % 2 else None)
        compare(got, exp, f"seed {seed} {kw} {subsets}")


@pytest.mark.parametrize("seed", range(6))
def test_multi_tile_contigs(oracle, seed):
    """Contigs spanning several 256-row warp tiles and several rounds of the CTA (the cross-warp
    insertion-chain carry, per-contig reductions), low-complexity so that long gaps and ties occur."""
    rng = random.Random(100 + seed)
    alphabet = [b"ACGT", b"AC", b"ACGT"][seed % 3]
    lens = [rng.randint(200, 2600) for _ in range(rng.randint(2, 5))] + [257, 256, 255][: seed % 3 + 1]
    contigs = [gen.rand_seq(rng, l, alphabet) for l in lens]
    reads = [gen.chimeric_read(rng, contigs, rng.randint(150, 500), rng.randint(1, 4), strands=True, wrap=True,
                               alphabet=alphabet) for _ in range(5)]
    kw = gen.fuzz_opts(seed)
    kw["double_strand"] = seed % 2 == 0
    if seed == 3:   # long insertions across tile boundaries: cheap gaps
        kw.update(gap_open=-1, gap_extend=0, mismatch_score=-6)
    got, exp = run_both(oracle, kw, contigs, reads, raw=(seed % 2 == 1),
                        tuning={"STITCH_CK_EVERY": "50", "STITCH_TRACK_WINDOW": "40"} if seed >= 3 else None)
    compare(got, exp, f"seed {seed} {kw}")


def test_config1_shape_downscaled(oracle):
    """BASELINE config 1/2 shapes, down-scaled so the oracle finishes in seconds: 20 plasmid-like
    contigs, chimeric noisy reads, default CLI scoring; single strand and -d -C."""
    rng = random.Random(20241)
    contigs = [gen.rand_seq(rng, rng.randint(700, 900)) for _ in range(20)]
    reads = [gen.chimeric_read(rng, contigs, 1000, rng.randint(3, 6)) for _ in range(6)]
    got, exp = run_both(oracle, {}, contigs, reads, raw=False)
    compare(got, exp, "config1")
    rng = random.Random(20242)
    reads = [gen.chimeric_read(rng, contigs, 1000, rng.randint(3, 6), strands=True, wrap=True) for _ in range(4)]
    got, exp = run_both(oracle, dict(double_strand=True, circular=True), contigs, reads, raw=False)
    compare(got, exp, "config2")
    # same batch split into chunks of 2 reads in flight: identical answers
    got2, _ = run_both(oracle, dict(double_strand=True, circular=True), contigs, reads, raw=False, max_inflight=2)
    compare(got2, exp, "config2 chunked")


def test_full_size_properties():
    """Config-1 sizes (10 kb reads vs 20 x ~8 kb contigs), no oracle: size-independent properties.
    A read that is an exact concatenation of contig substrings must score
    sum(len) + jumps * jump_score in local mode, its chain must validate, and the alignment of the
    read's reverse complement under --double-strand must have the same score."""
    import stitch_b200
    rng = random.Random(5)
    contigs = [gen.rand_seq(rng, rng.randint(7000, 9000)) for _ in range(20)]
    segs = []
    for _ in range(4):
        c = rng.randrange(20)
        s = rng.randrange(0, len(contigs[c]) - 2500)
        segs.append(contigs[c][s:s + 2500])
    read = b"".join(segs)
    named = [(f"c{k}", s) for k, s in enumerate(contigs)]
    al = gpu_aligners(dict(double_strand=True), named)
    chains = al.align_batch([read, gen.revcomp(read)])
    fwd, rev = chains[0][0], chains[1][0]
    assert fwd.score == 10000 - 3 * 10
    assert rev.score == fwd.score
    assert fwd.length == 10000 and fwd.ystart == 0 and fwd.yend == 10000
    fwd.validate(); rev.validate()
    assert sum(1 for k, _, _ in fwd.ops if k == 6) == 3
    st = al.stats()
    assert st.cells == 2 * 10000 * 2 * sum(len(c) for c in contigs)


@pytest.mark.parametrize("block", range(4))
def test_fuzz_packed_small(oracle, block):
    """Scorings inside the packed kernel's regime (dp_packed.h), tiny checkpoint spacing / wide tail so that
    small reads run through packed columns, the hand-over, wide re-filled units and the walk."""
    for seed in range(block * 50, block * 50 + 50):
        alphabet = [b"ACGT", b"AC", b"A", b"ACGTN"][seed % 4]
        contigs, reads = gen.fuzz_case(seed + 7000, max_contigs=5, max_len=90, max_read=70, alphabet=alphabet)
        kw = gen.fuzz_opts_packed(seed, 8)
        got, exp = run_both(oracle, kw, contigs, reads, raw=(seed % 3 != 0), tuning=cluster_tuning(seed, TUNING))
        compare(got, exp, f"seed {seed} {kw}")


@pytest.mark.parametrize("block", range(4))
def test_fuzz_packed_multi_tile(oracle, block):
    """Contigs spanning several 256-row tiles and all 16 warp chunks of the packed kernel (halo strips),
    reads long enough for several checkpoint blocks and for the column base to drift."""
    for seed in range(block * 6, block * 6 + 6):
        rng = random.Random(9000 + seed * 31)
        alphabet = [b"ACGT", b"AC", b"ACG"][seed % 3]
        lens = [rng.randint(1, 2600) for _ in range(rng.randint(1, 6))] + [257, 256, 255][: seed % 4]
        if seed % 5 == 0:
            lens.append(9000)          # 36 tiles: every warp chunk of the CTA starts inside this contig
        contigs = [gen.rand_seq(rng, l, alphabet) for l in lens]
        reads = [gen.chimeric_read(rng, contigs, rng.randint(20, 300), rng.randint(1, 4), strands=True,
                                   wrap=rng.random() < 0.5, noise=rng.random() < 0.8, alphabet=alphabet) for _ in range(4)]
        reads = [r if r else b"A" for r in reads]
        kw = gen.fuzz_opts_packed(seed + 100 * block, 8)
        got, exp = run_both(oracle, kw, contigs, reads, raw=(seed % 2 == 0),
                            tuning=cluster_tuning(seed + block, {"STITCH_CK_EVERY": "50", "STITCH_TRACK_WINDOW": "8"}))
        compare(got, exp, f"seed {seed} {kw}")


@pytest.mark.parametrize("mode", [0, 1])
def test_tail_restart_columns(oracle, mode):
    """The tail of the packed fill restarts from a regular checkpoint (every K columns) or from one of the late ones at
    n - 32, n - 64, n - 96 (kernels_packed.cuh): read lengths on both sides of every such column, with K = 64 so that late
    checkpoints coincide with regular ones for some lengths, exist only in part for short reads, and are the restart column for
    others; local mode (tracking threshold of the best-end walk, dp_core.h track_margin) and query-local mode (full margin)."""
    rng = random.Random(4242 + mode)
    contigs = [gen.rand_seq(rng, rng.randint(300, 700)) for _ in range(3)]
    lengths = [1, 20, 31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 129, 159, 160, 161, 191, 192, 193, 224, 250, 256, 257, 300]
    reads = []
    for n in lengths:
        r = gen.chimeric_read(rng, contigs, n, 1 + n // 120, strands=True, wrap=True, noise=True)
        r = (r + gen.rand_seq(rng, n))[:n]   # exactly n columns (the noise changes the length)
        reads.append(r)
    kw = dict(mode=mode, double_strand=True, circular=True)
    for tuning in ({"STITCH_CK_EVERY": "64"}, {"STITCH_CK_EVERY": "64", "STITCH_QUIET": "0"}, {"STITCH_CK_EVERY": "96", "STITCH_WALK_IN_KERNEL": "0"}):
        got, exp = run_both(oracle, kw, contigs, reads, raw=False, tuning=tuning)
        compare(got, exp, f"mode {mode} {tuning}")
        got, exp = run_both(oracle, kw, contigs, reads, raw=True, tuning=tuning)
        compare(got, exp, f"mode {mode} raw {tuning}")


def test_config3_config4_shapes_downscaled(oracle):
    """BASELINE config 3 (many contigs: the reference's limit of 256 contig-strands) and config 4 (reads several
    times longer than the contigs, many segments) down-scaled so the oracle finishes in seconds."""
    rng = random.Random(20243)
    contigs = [gen.rand_seq(rng, rng.randint(50, 100)) for _ in range(128)]      # 256 contig-strands
    reads = [gen.chimeric_read(rng, contigs, rng.randint(100, 250), rng.randint(3, 8), strands=True) for _ in range(4)]
    got, exp = run_both(oracle, dict(double_strand=True), contigs, reads, raw=False, tuning={"STITCH_CK_EVERY": "32"})
    compare(got, exp, "config3 slice")
    rng = random.Random(20244)
    contigs = [gen.rand_seq(rng, 400) for _ in range(25)]
    reads = [gen.chimeric_read(rng, contigs, rng.randint(2000, 3000), rng.randint(10, 30), strands=True) for _ in range(3)]
    got, exp = run_both(oracle, dict(double_strand=True), contigs, reads, raw=False)
    compare(got, exp, "config4")
    got, exp = run_both(oracle, dict(double_strand=True, suboptimal=True), contigs, reads[:1], raw=False)
    compare(got, exp, "config4 suboptimal")


def _exact_read(rng, contigs, nseg, seglen):
    segs = []
    for _ in range(nseg):
        c = rng.randrange(len(contigs))
        s = rng.randrange(0, len(contigs[c]) - seglen)
        piece = contigs[c][s:s + seglen]
        segs.append(gen.revcomp(piece) if rng.random() < 0.5 else piece)
    return b"".join(segs)


def test_config3_config4_full_size_properties():
    """Full-size shapes of configs 3 and 4 (no oracle): reads that are exact concatenations of contig substrings must
    score sum(len) + jumps * jump_score with one jump per junction, span the whole read, and validate.  Exercises 19
    length bits, 256 contig-strands / 2 M rows (7.8 k tiles), and the widened checkpoint spacing."""
    rng = random.Random(6)
    contigs = [gen.rand_seq(rng, rng.randint(5000, 10000)) for _ in range(128)]   # config 3 slice: 256 contig-strands
    named = [(f"c{k}", s) for k, s in enumerate(contigs)]
    read = _exact_read(rng, contigs, 5, 2400)
    al = gpu_aligners(dict(double_strand=True), named)
    a = al.align_batch([read])[0][0]
    assert (a.score, a.length, a.ystart, a.yend) == (12000 - 4 * 10, 12000, 0, 12000)
    a.validate()
    assert sum(1 for k, _, _ in a.ops if k == 6) == 4
    al.close()
    rng = random.Random(7)
    contigs = [gen.rand_seq(rng, 20000) for _ in range(50)]                        # config 4: 1 Mb panel, both strands
    named = [(f"c{k}", s) for k, s in enumerate(contigs)]
    read = _exact_read(rng, contigs, 12, 5000)                                     # 60 kb, 12 segments
    al = gpu_aligners(dict(double_strand=True), named)
    a = al.align_batch([read])[0][0]
    assert (a.score, a.length, a.ystart, a.yend) == (60000 - 11 * 10, 60000, 0, 60000)
    a.validate()
    assert sum(1 for k, _, _ in a.ops if k == 6) == 11
    st = al.stats()
    assert st.packed_cells == st.cells == 60000 * 2 * 50 * 20000
    al.close()


def test_sam_records_on_gpu(oracle):
    """Aligners::align + SamRecordFormatter::format through the C ABI on the GPU (stitch_format_sam) against the
    restated reference (oracle/ aligner + oracle/sam_oracle.py)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import sam_oracle
    import stitch_b200
    rng = random.Random(31)
    contigs = [gen.rand_seq(rng, rng.randint(300, 700)) for _ in range(5)]
    reads = [gen.chimeric_read(rng, contigs, 500, rng.randint(2, 5), strands=True, wrap=True) for _ in range(6)]
    named = [(f"ctg{k}", s) for k, s in enumerate(contigs)]
    headers = [f"r{k} extra" for k in range(len(reads))]
    quals = [bytes(rng.randrange(33, 74) for _ in r) for r in reads]
    for kw, so in ((dict(double_strand=True, circular=True), dict()),
                   (dict(double_strand=True, suboptimal=True, suboptimal_pct=0.0), dict(soft_clip=True, use_eq_and_x=True, pick_primary=1)),
                   (dict(double_strand=True, suboptimal=True, suboptimal_pct=0.0), dict(filter_secondary=True, filter_secondary_pct=50.0))):
        o = make_opts(**kw)
        exp_chains, _ = oracle.OracleAligners(o, named).batch(reads, raw=False)
        al = stitch_b200.Builder(**kw).build_aligners([stitch_b200.TargetSeq(n, s) for n, s in named], device=0)
        chains, sam = al.align_batch_sam(reads, headers, quals, so)
        al.close()
        compare(chains, exp_chains, f"sam {kw}")
        for r in range(len(reads)):
            exp = sam_oracle.format_sam(headers[r], reads[r], quals[r], exp_chains[r], [(n, len(s)) for n, s in named],
                                        (o.match_score, o.mismatch_score, o.gap_open, o.gap_extend), **so)
            assert sam[r] == exp, f"read {r} {kw} {so}"


@pytest.mark.parametrize("case", range(8))
def test_quiet_tiles_on_gpu(oracle, case):
    """Quiet tiles (dp_packed.h): with the default checkpoint spacing the bulk pass skips the warp tiles that are
    provably in the closed form "jump + substitution score" and re-materialises them when needed.  Same chains
    as the oracle, a good share of the tile-columns skipped, and the same chains again with STITCH_QUIET=0."""
    rng = random.Random(4400 + case)
    kw = [dict(double_strand=True, circular=True), dict(mode=1, double_strand=True), dict(mode=3, circular=True),
          dict(mode=2, double_strand=True, match_score=2, mismatch_score=-3, gap_open=-4, gap_extend=-3,
               jump_score_same_contig_and_strand=-8, jump_score_same_contig_opposite_strand=-9, jump_score_inter_contig=-11),
          dict(circular=True), dict(mode=3, double_strand=True, circular=True), dict(mode=1, circular=True, suboptimal=True),
          dict(mode=2, double_strand=True, circular=True, gap_extend=-3, jump_score_same_contig_and_strand=-6, jump_score_inter_contig=-12)][case]
    # lengths around the 256-row tile size: row m alone in the last tile (m = 256 k + 1), m a multiple of 256, a two-tile contig
    contigs = [gen.rand_seq(rng, rng.randint(2300, 3500)) for _ in range(3)] + [gen.rand_seq(rng, l) for l in (700, 2305, 2560, 300)]
    if case >= 4:   # a lower-complexity alphabet makes score / length ties frequent
        contigs = [gen.rand_seq(rng, len(c), b"ACG") if k % 2 else c for k, c in enumerate(contigs)]
    reads = [gen.chimeric_read(rng, contigs, rng.randint(700, 1100), rng.randint(2, 4), strands=bool(kw.get("double_strand")),
                               wrap=bool(kw.get("circular"))) for _ in range(5)]
    named = [(f"c{k}", s) for k, s in enumerate(contigs)]
    exp, _ = oracle.OracleAligners(make_opts(**kw), named).batch(reads, raw=False)
    # (few reads would otherwise get a cluster per read: no quiet tiles there)
    al = gpu_aligners(kw, named, tuning={"STITCH_CLUSTER": "1"})
    got = al.align_batch(reads)
    st = al.stats()
    al.close()
    compare(got, exp, f"quiet case {case}")
    assert st.tile_columns > 0 and st.quiet_tile_columns > 0.1 * st.tile_columns, (st.tile_columns, st.quiet_tile_columns)
    al = gpu_aligners(kw, named, tuning={"STITCH_QUIET": "0", "STITCH_CLUSTER": "1"})
    got0 = al.align_batch(reads)
    st0 = al.stats()
    al.close()
    compare(got0, exp, f"quiet off case {case}")
    assert st0.quiet_tile_columns == 0


def test_cli_on_gpu(oracle, tmp_path):
    """`stitch-b200 align` (stitch_b200/csrc/stitch_align_cli.cpp) end to end on the device: FASTQ + FASTA in, SAM text
    out, against the restated reference (oracle/ aligner + oracle/sam_oracle.py); records in input order."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle"))
    import sam_oracle
    cli = os.path.join(root, "stitch_b200", "stitch-b200")
    if not os.path.exists(cli):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", cli, os.path.join(root, "stitch_b200", "csrc", "stitch_align_cli.cpp"), "-ldl", "-lz"])
    rng = random.Random(77)
    contigs = [gen.rand_seq(rng, rng.randint(400, 900)) for _ in range(4)]
    reads = [gen.chimeric_read(rng, contigs, rng.randint(200, 600), rng.randint(1, 4), strands=True, wrap=True) for _ in range(10)]
    reads.insert(5, reads[4])
    named = [(f"ctg{k}", s) for k, s in enumerate(contigs)]
    heads = [f"r{k} x" for k in range(len(reads))]
    quals = [bytes(rng.randrange(35, 74) for _ in r) for r in reads]
    ref, fq = tmp_path / "ref.fa", tmp_path / "reads.fq"
    ref.write_text("".join(f">{n}\n{s.decode()}\n" for n, s in named))
    fq.write_text("".join(f"@{h}\n{r.decode()}\n+\n{q.decode()}\n" for h, r, q in zip(heads, reads, quals)))
    kw = dict(double_strand=True, circular=True)
    o = make_opts(**kw)
    exp_chains, _ = oracle.OracleAligners(o, named).batch(reads, raw=False)
    exp = []
    for r in range(len(reads)):
        exp += sam_oracle.format_sam(heads[r], reads[r], quals[r], exp_chains[r], [(n, len(s)) for n, s in named],
                                     (o.match_score, o.mismatch_score, o.gap_open, o.gap_extend))
    env = {k: v for k, v in os.environ.items() if k not in ("STITCH_B200_LIB", "STITCH_B200_PREFIX")}
    p = subprocess.run([cli, "align", "-f", str(fq), "-r", str(ref), "-d", "-C", "--sam", "--batch", "3"], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()
    got = [l for l in p.stdout.decode().splitlines() if not l.startswith("@")]
    assert got == exp


def _oracle_threads(bytes_per_read):
    try:
        avail = int(next(l for l in open("/proc/meminfo") if l.startswith("MemAvailable")).split()[1]) * 1024
    except Exception:
        avail = 32 << 30
    return int(max(1, min(os.cpu_count() or 1, 8, avail * 0.6 // max(1, bytes_per_read))))


def _full_size_case(oracle, config, n_batch, n_check, read_len=None, pick=None):
    """`n_batch` full-size reads of a BASELINE config through the GPU with PRODUCTION defaults (no tuning knob set: default
    checkpoint spacing, one CTA per read, quiet tiles), then `n_check` of them again with each read restricted
    (subset_words) to the contig-strands its segments were drawn from plus two decoys, bit for bit against the oracle on
    the same subsets (the oracle's 16-byte cells make all contigs unaffordable: a 10 kb read against 7 strands is 9 GB)."""
    import stitch_b200
    from stitch_b200 import synth
    truth = []
    kw, named, reads = synth.config(config, n_batch, read_len, truth=truth)
    ns = len(named) * (2 if kw.get("double_strand") else 1)
    rng = random.Random(config)
    al = gpu_aligners(kw, named)          # production defaults
    full = al.align_batch(reads)
    st = al.stats()
    assert st.packed_cells > 0 and st.quiet_tile_columns > 0.3 * st.tile_columns, (st.tile_columns, st.quiet_tile_columns)
    subsets = []
    for r in range(n_batch):
        # (a spurious short hit of the read's random tail may add a contig the segments were not drawn from)
        used = {full[r][0].start_contig_idx, full[r][0].end_contig_idx} | {a for k, a, _ in full[r][0].ops if k == 6}
        mine = set(truth[r]) | used
        decoys = rng.sample([c for c in range(ns) if c not in mine], 2)
        subsets.append(sorted(mine | set(decoys)))
    check = sorted(range(n_batch), key=lambda r: (len(subsets[r]), r))[:n_check] if pick is None else pick
    sub = al.align_batch(reads, subsets)
    st2 = al.stats()
    assert st2.quiet_tile_columns > 0.1 * st2.tile_columns, (st2.tile_columns, st2.quiet_tile_columns)
    al.close()
    for r in range(n_batch):   # the subset holds every contig of the read: same optimum as against all contigs
        assert full[r][0].score == sub[r][0].score, f"config {config} read {r}: {full[r][0].score} vs {sub[r][0].score}"
    rows = max(sum(len(named[c % len(named)][1]) + 1 for c in subsets[r]) for r in check)
    per_read = 16 * rows * (max(len(reads[r]) for r in check) + 1)
    oracle.set_checker_layout(True)
    try:
        exp, info = oracle.OracleAligners(make_opts(**kw), named).batch([reads[r] for r in check], subsets=[subsets[r] for r in check],
                                                                         raw=False, threads=_oracle_threads(per_read))
    finally:
        oracle.set_checker_layout(False)
    compare([sub[r] for r in check], exp, f"config {config} full size")
    return info


def test_full_size_config2_vs_oracle(oracle):
    """BASELINE config 2 at the size the headline is quoted on: noisy 10 kb reads, --double-strand --circular (origin
    re-alignment fills included), 64 reads so that the backend runs one CTA per read with quiet tiles."""
    _full_size_case(oracle, 2, 64, 6)


def test_full_size_config1_vs_oracle(oracle):
    _full_size_case(oracle, 1, 64, 4)


def test_full_size_config3_slice_vs_oracle(oracle):
    """A config-3-slice read of 20 kb (256 contig-strands in the table, 17 length bits)."""
    _full_size_case(oracle, 3, 60, 1, read_len=20000)


def test_many_reads_on_the_wide_path(oracle):
    """More reads than 2 x SM count on the wide kernels (scorings outside the packed regime, here gap_extend = -1 with the
    default jump score: insertion-chain reach 16 > 8): every CTA of the wide fill's grid owns its own rolling state."""
    rng = random.Random(811)
    contigs = [gen.rand_seq(rng, rng.randint(60, 300)) for _ in range(3)]
    reads = [gen.chimeric_read(rng, contigs, rng.randint(30, 90), rng.randint(1, 3), strands=True) for _ in range(400)]
    reads = [r if r else b"A" for r in reads]
    for kw, tuning in ((dict(double_strand=True, gap_extend=-1), None), (dict(double_strand=True, mode=1), {"STITCH_PACKED": "0"})):
        got, exp = run_both(oracle, kw, contigs, reads, raw=False, tuning=tuning)
        compare(got, exp, f"wide path {kw}")


def test_config4_many_long_reads():
    """BASELINE config 4 at its stated size on the in-kernel walk path (>= 56 reads, so one CTA per read): 60 reads of 70-100 kb
    against 50 x 20 kb contigs, both strands.  The checkpoint spacing widens past 2 k columns; the walk phase's shared-memory
    staging is sized with that spacing (it used to exceed the 227 KB limit).  Exact-concatenation reads: score, span, jumps."""
    rng = random.Random(44)
    contigs = [gen.rand_seq(rng, 20000) for _ in range(50)]
    named = [(f"c{k}", s) for k, s in enumerate(contigs)]
    nseg = [rng.randint(10, 14) for _ in range(60)]
    reads = [_exact_read(rng, contigs, k, rng.randint(6000, 7000)) for k in nseg]
    al = gpu_aligners(dict(double_strand=True), named)
    res = al.align_batch(reads)
    st = al.stats()
    al.close()
    for k, read, chains in zip(nseg, reads, res):
        a = chains[0]
        assert (a.score, a.length, a.ystart, a.yend) == (len(read) - (k - 1) * 10, len(read), 0, len(read))
        a.validate()
    assert st.packed_cells == st.cells


def test_contig_longer_than_the_staging_area(oracle):
    """A 300 kb contig: its bases do not fit the walk's shared-memory staging area and are read from global memory
    instead (packed path), and the wide path's walk never stages them."""
    rng = random.Random(45)
    contigs = [gen.rand_seq(rng, 300000)] + [gen.rand_seq(rng, rng.randint(500, 900)) for _ in range(2)]
    named = [(f"c{k}", s) for k, s in enumerate(contigs)]
    reads = []
    for _ in range(4):
        s = rng.randrange(0, 300000 - 700)
        piece = contigs[0][s:s + 700]
        other = contigs[1][100:400]
        reads.append(piece + (gen.revcomp(other) if rng.random() < 0.5 else other))
    for tuning in (None, {"STITCH_WALK_IN_KERNEL": "0"}, {"STITCH_PACKED": "0"}):
        al = gpu_aligners(dict(double_strand=True), named, tuning=tuning)
        res = al.align_batch(reads)
        al.close()
        for read, chains in zip(reads, res):
            a = chains[0]
            assert (a.score, a.length, a.ystart, a.yend) == (1000 - 10, 1000, 0, 1000), tuning
            a.validate()
    # one noisy read against the oracle (16 B x 300 kb x 400 columns = 2 GB)
    noisy = [gen.noisy(rng, reads[0][:400])]
    got, exp = run_both(oracle, dict(double_strand=False), contigs, noisy, raw=False)
    compare(got, exp, "long contig")


class _GpuPre:
    """The shape tests/test_prealign.py's checks expect, over the CUDA library."""
    def __init__(self, kw, named):
        self.al = gpu_aligners(kw, named)
    def batch(self, reads):
        out = self.al.align_batch(reads)
        self.last_prealign_scores = self.al.last_prealign_scores
        return out
    def prealign_batch(self, reads):
        return self.al.prealign_batch(reads)
    def close(self):
        self.al.close()


def test_prealign_on_gpu(oracle):
    """Pre-alignment contig selection on the CUDA kernel (kernels_prealign.cuh): the same selections and scores as the
    sequential statement of prealign_core.h in the CPU emulator, chains equal to the oracle's on the selected subsets, a
    600-strand contig table, direct (k <= 12) and hashed (k = 15) buckets."""
    import emul_lib
    import test_prealign as tp
    for seed, kw, shape in ((11, dict(double_strand=True, pre_align=True, kmer_size=8, band_width=20, pre_align_min_score=40), (12, 300, 600, 6, 400, 3)),
                            (13, dict(double_strand=True, pre_align=True, kmer_size=10, band_width=30, pre_align_min_score=50), (300, 120, 260, 5, 360, 3)),
                            (14, dict(double_strand=True, circular=True, pre_align=True, kmer_size=15, band_width=50, pre_align_min_score=80), (40, 400, 900, 5, 900, 3))):
        contigs, reads, truth = tp.make_case(seed, *shape)
        sel, best = tp.check_subset_mode(oracle, _GpuPre, kw, contigs, reads, truth)
        e = emul_lib.EmulAligners(make_opts(**kw), [(f"c{k}", s) for k, s in enumerate(contigs)], strip=8)
        assert (sel, best) == e.prealign_batch(reads), kw
        e.close()
    # reads that reach the score nowhere come back unmapped
    rng = random.Random(5)
    contigs = [gen.rand_seq(rng, 400) for _ in range(4)]
    al = gpu_aligners(dict(double_strand=True, pre_align=True), [(f"c{k}", s) for k, s in enumerate(contigs)])
    chains = al.align_batch([contigs[2][50:350], gen.rand_seq(rng, 300), contigs[1][10:60]])
    assert [len(c) for c in chains] == [1, 0, 0] and al.last_prealign_scores == [300, None, None]
    al.close()


def test_config3_full_table_with_prealign(oracle):
    """BASELINE config 3 at its stated scale: 2 000 contigs of 5-10 kb x 2 strands = 4 000 contig-strands (15 Mb, 30 M indexed
    k-mers), reads of 5-20 kb with 3-8 segments, `-d -p -x`.  Every contig-strand a read was drawn from is selected, the
    selection stays far below 256 strands, and the chains of three reads equal the oracle's on the same subsets."""
    import numpy as np
    import test_prealign as tp
    from stitch_b200 import synth
    rng = np.random.default_rng(20243)
    contigs = [c.tobytes() for c in synth.make_contigs(rng, 2000, 5000, 10000)]
    reads, truth = [], []
    for _ in range(48):
        t = []
        reads.append(synth.make_read(rng, [np.frombuffer(c, dtype=np.uint8) for c in contigs], int(rng.integers(5000, 20001)), 3, 8,
                                     strands=True, truth=t).tobytes())
        truth.append(sorted(set(t)))
    kw = dict(double_strand=True, pre_align=True)
    named = [(f"c{k}", s) for k, s in enumerate(contigs)]
    al = gpu_aligners(kw, named)
    got = al.align_batch(reads)
    pre = al.last_prealign_scores
    st = al.stats()
    sel, best = al.prealign_batch(reads)
    al.close()
    assert st.prealign_reads == 48 and st.packed_cells == st.cells
    for r in range(len(reads)):
        # a segment shorter than ~150 bases may stay below the minimum score of 100, like in the reference
        assert len(sel[r]) <= 32 and pre[r] == best[r] and len(got[r]) == 1
        got[r][0].validate()
    long_truth = sum(1 for r in range(len(reads)) for c in truth[r] if c in sel[r])
    assert long_truth >= 0.9 * sum(len(t) for t in truth), (long_truth, sum(len(t) for t in truth))
    for r in sorted(range(len(reads)), key=lambda r: (len(reads[r]) * len(sel[r]), r))[:3]:
        oracle.set_checker_layout(True)
        try:
            exp = tp.reduced_oracle(oracle, kw, contigs, reads[r], sel[r])
        finally:
            oracle.set_checker_layout(False)
        assert [a.key() for a in got[r]] == [a.key() for a in exp], r
