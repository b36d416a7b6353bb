"""CPU tier: the `stitch align` front-end (stitch_b200/csrc/stitch_align_cli.cpp) - option parsing, FASTA/FASTQ(.gz)
readers, the align-identical-runs-once rule, output order, SAM text and BAM/BGZF encoding.  The binary binds the
C ABI at run time; here it is pointed at the CPU emulator of the kernels (tests/emul, TEST INFRASTRUCTURE), so the
records must equal what the emulator's own batch_sam returns (which test_sam_layer.py checks against the
restated reference), and the BAM stream must decode to the same records."""
import gzip
import os
import random
import struct
import subprocess
import zlib

import pytest

import gen
from stitch_b200._abi import make_opts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "stitch_b200", "stitch-b200")
SRC = os.path.join(ROOT, "stitch_b200", "csrc", "stitch_align_cli.cpp")


@pytest.fixture(scope="module")
def cli():
    import emul_lib
    emul_lib.build()
    if not os.path.exists(CLI) or os.path.getmtime(CLI) < os.path.getmtime(SRC):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", CLI, SRC, "-ldl", "-lz"])
    env = dict(os.environ, STITCH_B200_LIB=os.path.join(ROOT, "tests", "emul", "libemul_s8.so"), STITCH_B200_PREFIX="emul_")

    def run(args):
        p = subprocess.run([CLI, "align"] + args, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert p.returncode == 0, p.stderr.decode()
        return p.stdout
    return run


def bgzf_decompress(data):
    out, p = bytearray(), 0
    while p < len(data):
        assert data[p:p + 4] == b"\x1f\x8b\x08\x04"
        xlen = struct.unpack_from("<H", data, p + 10)[0]
        assert data[p + 12:p + 16] == b"BC\x02\x00"
        bsize = struct.unpack_from("<H", data, p + 16)[0] + 1
        cdata = data[p + 12 + xlen:p + bsize - 8]
        crc, isize = struct.unpack_from("<II", data, p + bsize - 8)
        block = zlib.decompress(cdata, -15) if cdata else b""
        assert len(block) == isize and zlib.crc32(block) == crc
        out += block
        p += bsize
    return bytes(out), data[-28:]


def bam_to_sam(raw):
    """Minimal BAM decoder -> (header text, reference names, SAM lines)."""
    assert raw[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<I", raw, 4)[0]
    text = raw[8:8 + l_text].decode()
    p = 8 + l_text
    n_ref = struct.unpack_from("<I", raw, p)[0]; p += 4
    refs = []
    for _ in range(n_ref):
        l = struct.unpack_from("<I", raw, p)[0]; p += 4
        refs.append((raw[p:p + l - 1].decode(), struct.unpack_from("<I", raw, p + l)[0])); p += l + 4
    lines = []
    while p < len(raw):
        bs = struct.unpack_from("<I", raw, p)[0]; p += 4
        rec = raw[p:p + bs]; p += bs
        rid, pos, l_name, mapq, _bin, n_cig, flag, l_seq, nrid, npos, tlen = struct.unpack_from("<iiBBHHHIiii", rec, 0)
        q = 32
        name = rec[q:q + l_name - 1].decode(); q += l_name
        cig = ""
        for _ in range(n_cig):
            v = struct.unpack_from("<I", rec, q)[0]; q += 4
            cig += f"{v >> 4}{'MIDNSHP=X'[v & 15]}"
        seq = "".join("=ACMGRSVTWYHKDBN"[(rec[q + k // 2] >> (4 if k % 2 == 0 else 0)) & 15] for k in range(l_seq)); q += (l_seq + 1) // 2
        qual = rec[q:q + l_seq]; q += l_seq
        qual = "*" if l_seq == 0 or qual[0] == 0xff else "".join(chr(c + 33) for c in qual)
        tags = []
        while q < len(rec):
            tag, ty = rec[q:q + 2].decode(), chr(rec[q + 2]); q += 3
            if ty in "cCsSiI":
                fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}[ty]
                tags.append(f"{tag}:i:{struct.unpack_from(fmt, rec, q)[0]}"); q += struct.calcsize(fmt)
            elif ty == "Z":
                e = rec.index(0, q); tags.append(f"{tag}:Z:{rec[q:e].decode()}"); q = e + 1
            elif ty == "A":
                tags.append(f"{tag}:A:{chr(rec[q])}"); q += 1
            else:
                raise AssertionError(ty)
        rn = lambda i, same: "*" if i < 0 else ("=" if same and i == rid else refs[i][0])
        lines.append("\t".join([name, str(flag), rn(rid, False), str(pos + 1), str(mapq), cig or "*", rn(nrid, True), str(npos + 1), str(tlen),
                                seq or "*", qual] + tags))
    return text, refs, lines


def test_cli_sam_and_bam(cli, tmp_path):
    import emul_lib
    rng = random.Random(31)
    contigs = [gen.rand_seq(rng, rng.randint(150, 400)) for _ in range(3)]
    reads = [gen.chimeric_read(rng, contigs, rng.randint(60, 200), rng.randint(1, 3), strands=True, wrap=True) for _ in range(9)]
    reads.insert(3, reads[2]); reads.insert(4, reads[2].lower())          # a run of identical sequences: aligned once
    reads.append(gen.rand_seq(rng, 40))                                    # unrelated to the contigs: a short local hit somewhere
    heads = [f"read{k} extra words {k}" for k in range(len(reads))]
    quals = ["".join(chr(33 + rng.randint(2, 40)) for _ in r) for r in reads]
    ref = tmp_path / "ref.fa"
    with open(ref, "w") as f:
        for k, c in enumerate(contigs):
            s = c.decode()
            s = s.lower() if k == 1 else s
            f.write(f">ctg{k} some description\n" + "\n".join(s[i:i + 70] for i in range(0, len(s), 70)) + "\n")
    fq = tmp_path / "reads.fq.gz"
    with gzip.open(fq, "wt") as f:
        for h, r, q in zip(heads, reads, quals):
            f.write(f"@{h}\n{r.decode()}\n+\n{q}\n")
    fa = tmp_path / "reads.fa"
    with open(fa, "w") as f:
        for h, r in zip(heads, reads):
            f.write(f">{h}\n{r.decode()}\n")
    for extra, kw, sam_opts in (
            ([], {}, None),
            (["-d", "-C", "--circular-slop", "5", "-S", "-X"], dict(double_strand=True, circular=True, circular_slop=5), dict(soft_clip=True, use_eq_and_x=True)),
            (["-m", "query-local", "-d", "--suboptimal", "--suboptimal-pct=50", "-P", "score", "--filter-secondary", "--filter-secondary-pct", "40",
              "-A", "2", "-B", "-3", "-O", "-4", "-E", "-3", "-J", "-9", "--jump-score-inter-contig", "-11"],
             dict(mode=1, double_strand=True, suboptimal=True, suboptimal_pct=50.0, match_score=2, mismatch_score=-3, gap_open=-4, gap_extend=-3,
                  default_jump_score=-9, jump_score_inter_contig=-11), dict(pick_primary=1, filter_secondary=True, filter_secondary_pct=40.0))):
        named = [(f"ctg{k}", c) for k, c in enumerate(contigs)]
        e = emul_lib.EmulAligners(make_opts(**kw), named, strip=8)
        _, exp = e.batch_sam(reads, heads, quals=[q.encode() for q in quals], sam_opts=sam_opts)
        e.close()
        exp_lines = [l for per_read in exp for l in per_read]
        # SAM text, reads from gzip FASTQ, batches of 4 records (runs of identical sequences stay together)
        out = cli(["-f", str(fq), "-r", str(ref), "--sam", "--batch", "4"] + extra).decode().splitlines()
        hdr = [l for l in out if l.startswith("@")]
        assert hdr[0].startswith("@HD") and [l.split("\t")[1] for l in hdr if l.startswith("@SQ")] == ["SN:ctg0", "SN:ctg1", "SN:ctg2"]
        assert hdr[-1].startswith("@PG\tID:stitch")
        assert [l for l in out if not l.startswith("@")] == exp_lines
        # BAM on stdout (the reference's output), decoded back
        raw, tail = bgzf_decompress(cli(["-f", str(fq), "-r", str(ref), "-c", "1"] + extra))
        assert tail == bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])
        text, refs, lines = bam_to_sam(raw)
        assert refs == [(f"ctg{k}", len(c)) for k, c in enumerate(contigs)] and text.startswith("@HD")
        # (BAM's 4-bit base codes carry no case: the soft-masked read comes back upper-cased)
        assert lines == ["\t".join(f.upper() if k == 9 else f for k, f in enumerate(l.split("\t"))) for l in exp_lines]
        # FASTA reads: no qualities
        e = emul_lib.EmulAligners(make_opts(**kw), named, strip=8)
        _, exp = e.batch_sam(reads, heads, quals=None, sam_opts=sam_opts)
        e.close()
        out = cli(["-a", str(fa), "-r", str(ref), "--sam"] + extra).decode().splitlines()
        assert [l for l in out if not l.startswith("@")] == [l for per_read in exp for l in per_read]


def test_cli_rejects_what_it_cannot_do(cli, tmp_path):
    p = subprocess.run([CLI, "align", "-r", "x.fa"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0 and b"exactly one of --reads-fastq or --reads-fasta" in p.stderr
    p = subprocess.run([CLI, "align", "-f", "a.fq", "-r", "x.fa", "--bogus"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0 and b"unknown option --bogus" in p.stderr


def test_cli_pre_align(cli, tmp_path):
    """-p -k -w -s -x (align.rs:119-146): reads below the pre-alignment score come out as unmapped records without `xs`, the others
    carry the score in `xs`; same records as the library call with the same options."""
    import emul_lib
    rng = random.Random(91)
    contigs = [gen.rand_seq(rng, rng.randint(300, 500)) for _ in range(6)]
    reads = [gen.chimeric_read(rng, contigs, rng.randint(150, 300), rng.randint(1, 2), strands=True) for _ in range(5)]
    reads.insert(2, gen.rand_seq(rng, 120))          # matches nothing
    heads = [f"q{k}" for k in range(len(reads))]
    ref, fa = tmp_path / "ref.fa", tmp_path / "reads.fa"
    ref.write_text("".join(f">c{k}\n{c.decode()}\n" for k, c in enumerate(contigs)))
    fa.write_text("".join(f">{h}\n{r.decode()}\n" for h, r in zip(heads, reads)))
    for x in ("true", "false"):
        kw = dict(double_strand=True, pre_align=True, kmer_size=9, band_width=25, pre_align_min_score=45, pre_align_subset_contigs=(x == "true"))
        e = emul_lib.EmulAligners(make_opts(**kw), [(f"c{k}", c) for k, c in enumerate(contigs)], strip=8)
        _, exp = e.batch_sam(reads, heads, None, None)
        e.close()
        out = cli(["-a", str(fa), "-r", str(ref), "-d", "-p", "-k", "9", "-w", "25", "-s", "45", "-x", x, "--sam", "--batch", "2"]).decode().splitlines()
        got = [l for l in out if not l.startswith("@")]
        assert got == [l for per_read in exp for l in per_read]
        assert got[2 if x == "true" else 2].split("\t")[0] != "" and any(l.split("\t")[1] == "4" and "xs:i:" not in l for l in got)
        assert any("\txs:i:" in l for l in got)


def test_cli_multi_context_order(cli, tmp_path):
    """--gpus N: one context and one host thread per device, batches dealt round-robin; the records still come out in
    input order and equal the single-context output."""
    rng = random.Random(57)
    contigs = [gen.rand_seq(rng, rng.randint(150, 300)) for _ in range(3)]
    reads = [gen.chimeric_read(rng, contigs, rng.randint(50, 150), rng.randint(1, 3), strands=True, wrap=True) for _ in range(17)]
    ref, fa = tmp_path / "ref.fa", tmp_path / "reads.fa"
    ref.write_text("".join(f">c{k}\n{c.decode()}\n" for k, c in enumerate(contigs)))
    fa.write_text("".join(f">q{k}\n{r.decode()}\n" for k, r in enumerate(reads)))
    one = [l for l in cli(["-a", str(fa), "-r", str(ref), "-d", "--sam"]).decode().splitlines() if not l.startswith("@PG")]
    many = [l for l in cli(["-a", str(fa), "-r", str(ref), "-d", "--sam", "--gpus", "3", "--batch", "2"]).decode().splitlines() if not l.startswith("@PG")]
    assert one == many
    names = [l.split("\t")[0] for l in one if not l.startswith("@")]
    assert [n for k, n in enumerate(names) if k == 0 or names[k - 1] != n] == [f"q{k}" for k in range(len(reads))]


def test_cli_timing_line(cli, tmp_path):
    """STITCH_CLI_TIMING=1: one JSON line on stderr accounts for the stages of the pipeline (what `bench.py --via-cli` reads);
    without the variable stderr carries no such line and stdout is the same either way."""
    import json
    rng = random.Random(91)
    contigs = [gen.rand_seq(rng, rng.randint(150, 300)) for _ in range(2)]
    reads = [gen.chimeric_read(rng, contigs, rng.randint(50, 150), 2, strands=True, wrap=False) for _ in range(9)]
    ref, fa = tmp_path / "ref.fa", tmp_path / "reads.fa"
    ref.write_text("".join(f">c{k}\n{c.decode()}\n" for k, c in enumerate(contigs)))
    fa.write_text("".join(f">q{k}\n{r.decode()}\n" for k, r in enumerate(reads)))
    env = dict(os.environ, STITCH_B200_LIB=os.path.join(ROOT, "tests", "emul", "libemul_s8.so"), STITCH_B200_PREFIX="emul_")
    args = [CLI, "align", "-a", str(fa), "-r", str(ref), "-d", "--sam", "--gpus", "2", "--batch", "4"]
    quiet = subprocess.run(args, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    timed = subprocess.run(args, env=dict(env, STITCH_CLI_TIMING="1"), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert quiet.returncode == 0 and timed.returncode == 0, timed.stderr.decode()
    assert quiet.stdout == timed.stdout
    assert b"stitch-b200 timing:" not in quiet.stderr
    lines = [l for l in timed.stderr.decode().splitlines() if l.startswith("stitch-b200 timing: ")]
    assert len(lines) == 1
    t = json.loads(lines[0][len("stitch-b200 timing: "):])
    assert t["reads"] == len(reads) and t["gpus"] == 2
    assert 0 <= t["contexts_ready_s"] <= t["first_align_s"] <= t["last_align_s"] <= t["wall_s"]
    assert t["align_span_s"] > 0 and t["reads_per_s_in_align_span"] > 0 and t["format_threads"] >= 1
