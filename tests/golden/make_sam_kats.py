"""Writes tests/golden/sam_kats.json: HAND-DERIVED known-answer records for the SAM record layer.

The reference holds no test of SubAlignmentBuilder / SamRecordFormatter, so these vectors were
derived by hand from the reference source, statement by statement, and typed in below (they were NOT produced by
running oracle/sam_oracle.py or host_sam.hpp; both are checked against them by tests/test_sam_kats.py).  Each case
names the quirk it pins and the reference lines the derivation follows (LIB = fg-stitch-lib/src/align).

Operations are (kind, a, b) with kind 0 Match, 1 Subst, 2 Del (consumes a READ base), 3 Ins (consumes a CONTIG base),
6 Xjump(contig a, offset b), 7 Yjump(len a) - constants.rs:20-29, 61-84.
"""
import json
import os

M, X, D, I, XJ, YJ = 0, 1, 2, 3, 6, 7
TAB = "\t"


def rec(*fields):
    return TAB.join(str(f) for f in fields)


CASES = []

# ---- K1 ------------------------------------------------------------------------------------------------------------------
# Q13 (sub_alignment.rs:188-206): num_edits is incremented BEFORE the previous operation is flushed, so the Subst that is
# the first operation after the Xjump is credited to the sub-alignment the Xjump closes (NM 1 / 0 instead of 0 / 1).
# With use_eq_and_x = false cmp_op merges Subst and Match runs (:37-46) and add_op scores the merged run with the kind of
# its LAST operation (:195-206 `last = op`): the 1X4= run scores 5 x match = 5.
# Reverse strand (contig_idx >= #targets, mod.rs:760-761): SEQ = reverse_complement of the read AS GIVEN (mixed case kept,
# mod.rs:630, 790-798; dna.rs keeps case), QUAL reversed, POS = target_len - target_end + 1 (mod.rs:857-864), clips
# swapped (mod.rs:829-846).  Primary = longest query span (mod.rs:699-706).
CASES.append(dict(
    name="K1_nm_credit_reverse_strand",
    cites="sub_alignment.rs:37-46,48-103,169-241; mod.rs:699-706,756-809,822-864,936-966",
    targets=[["chrA", 20], ["chrB", 16]], double_strand=True, scoring=[1, -4, -6, -2],
    header="r1 first read", bases="ACGTTGcaGGAT", quals="ABCDEFGHIJKL", pre_align_score=None, sam_opts={},
    chains=[dict(score=-3, xstart=3, xend=9, ystart=0, yend=12, xlen=16, ylen=12, start_contig_idx=0, end_contig_idx=3, length=12,
                 ops=[[M, 7, 0], [XJ, 3, 4], [X, 1, 0], [M, 4, 0]])],
    expect=[
        rec("r1", 0, "chrA", 4, 60, "7M5S", "*", 0, 0, "ACGTTGcaGGAT", "ABCDEFGHIJKL", "qs:i:0", "qe:i:7", "ts:i:3", "te:i:10", "as:i:-3",
            "si:i:0", "sc:Z:7M", "cl:i:2", "ci:i:0", "cn:i:1", "AS:i:7", "NM:i:1", "SA:Z:chrA,4,+,7M5S,60,1;chrB,8,-,5M7S,60,0"),
        rec("r1", 2064, "chrB", 8, 60, "5M7S", "*", 0, 0, "ATCCtgCAACGT", "LKJIHGFEDCBA", "qs:i:7", "qe:i:12", "ts:i:4", "te:i:9", "as:i:-3",
            "si:i:1", "sc:Z:5M", "cl:i:2", "ci:i:0", "cn:i:1", "AS:i:5", "NM:i:0", "SA:Z:chrA,4,+,7M5S,60,1;chrB,8,-,5M7S,60,0"),
    ]))

# ---- K2 ------------------------------------------------------------------------------------------------------------------
# use_eq_and_x = true: '=' / 'X' elements, every run scored by its own kind.  Q2: Ins consumes a contig base and Del a read
# base (constants.rs:61-84), swapped into SAM D / I by swap_cigar (sub_alignment.rs:157-167).  SA rotation: the reference
# calls sa_strings.rotate_right(primary_sub_idx) (mod.rs:956): with primary index 1 of 3 the LAST string moves to the
# front ([A, B, C] -> [C, A, B]), i.e. the primary is NOT first - reproduced as is.  No qualities: QUAL '*'.
CASES.append(dict(
    name="K2_eq_x_indel_swap_sa_rotation",
    cites="sub_alignment.rs:48-103,157-167,169-241; mod.rs:699-706,822-846,936-966",
    targets=[["chrA", 30]], double_strand=False, scoring=[1, -4, -6, -2],
    header="r2", bases="AAAAACCCCCGGGGGTTTTT", quals=None, pre_align_score=None, sam_opts=dict(use_eq_and_x=True),
    chains=[dict(score=1, xstart=2, xend=2, ystart=1, yend=15, xlen=30, ylen=20, start_contig_idx=0, end_contig_idx=0, length=16,
                 ops=[[M, 3, 0], [XJ, 0, 10], [M, 2, 0], [X, 1, 0], [I, 2, 0], [M, 4, 0], [D, 1, 0], [M, 1, 0], [XJ, 0, 0], [M, 2, 0]])],
    expect=[
        rec("r2", 2048, "chrA", 3, 60, "1S3=16S", "*", 0, 0, "AAAAACCCCCGGGGGTTTTT", "*", "qs:i:1", "qe:i:4", "ts:i:2", "te:i:5", "as:i:1",
            "si:i:0", "sc:Z:3=", "cl:i:3", "ci:i:0", "cn:i:1", "AS:i:3", "NM:i:0",
            "SA:Z:chrA,1,+,13S2=5S,60,0;chrA,3,+,1S3=16S,60,0;chrA,11,+,4S2=1X2D4=1I1=7S,60,4"),
        rec("r2", 0, "chrA", 11, 60, "4S2=1X2D4=1I1=7S", "*", 0, 0, "AAAAACCCCCGGGGGTTTTT", "*", "qs:i:4", "qe:i:13", "ts:i:10", "te:i:20", "as:i:1",
            "si:i:1", "sc:Z:2=1X2D4=1I1=", "cl:i:3", "ci:i:0", "cn:i:1", "AS:i:-15", "NM:i:4",
            "SA:Z:chrA,1,+,13S2=5S,60,0;chrA,3,+,1S3=16S,60,0;chrA,11,+,4S2=1X2D4=1I1=7S,60,4"),
        rec("r2", 2048, "chrA", 1, 60, "13S2=5S", "*", 0, 0, "AAAAACCCCCGGGGGTTTTT", "*", "qs:i:13", "qe:i:15", "ts:i:0", "te:i:2", "as:i:1",
            "si:i:2", "sc:Z:2=", "cl:i:3", "ci:i:0", "cn:i:1", "AS:i:2", "NM:i:0",
            "SA:Z:chrA,1,+,13S2=5S,60,0;chrA,3,+,1S3=16S,60,0;chrA,11,+,4S2=1X2D4=1I1=7S,60,4"),
    ]))

# ---- K3 ------------------------------------------------------------------------------------------------------------------
# Two chains, default hard clipping.  Secondary chain (chain_idx > 0): flag 0x100, MAPQ 0 (mod.rs:868), SEQ / QUAL sliced to
# the sub-alignment (mod.rs:782-789, 799-809), clip operation H (mod.rs:822-826).  Q14: the FORWARD hard-clipped secondary
# record gets its CIGAR REVERSED (mod.rs:787-788: `sub.cigar.iter().rev()` in the (true, true) arm), and the `sc` tag is the
# string of that reversed CIGAR (cigar_str, mod.rs:811, 907-910).  xs = the best secondary chain score, on every record of every chain
# (mod.rs:678-687, 893-898).
K3_CHAINS = [
    dict(score=10, xstart=10, xend=20, ystart=0, yend=10, xlen=40, ylen=10, start_contig_idx=0, end_contig_idx=0, length=10,
         ops=[[M, 10, 0]]),
    dict(score=3, xstart=5, xend=9, ystart=2, yend=9, xlen=40, ylen=10, start_contig_idx=1, end_contig_idx=2, length=8,
         ops=[[M, 2, 0], [I, 1, 0], [M, 3, 0], [XJ, 2, 7], [M, 2, 0]]),
]
CASES.append(dict(
    name="K3_secondary_hard_clip_reversed_cigar",
    cites="mod.rs:678-687,756-809,822-868,893-898; sub_alignment.rs:157-167",
    targets=[["chrA", 40], ["chrB", 30]], double_strand=True, scoring=[1, -4, -6, -2],
    header="r3", bases="ACCGTTAGCA", quals="0123456789", pre_align_score=None, sam_opts={},
    chains=K3_CHAINS,
    expect=[
        rec("r3", 0, "chrA", 11, 60, "10M", "*", 0, 0, "ACCGTTAGCA", "0123456789", "qs:i:0", "qe:i:10", "ts:i:10", "te:i:20", "as:i:10", "xs:i:3",
            "si:i:0", "sc:Z:10M", "cl:i:1", "ci:i:0", "cn:i:2", "AS:i:10", "NM:i:0", "SA:Z:chrA,11,+,10M,60,0"),
        rec("r3", 256, "chrB", 6, 0, "2H3M1D2M3H", "*", 0, 0, "CGTTA", "23456", "qs:i:2", "qe:i:7", "ts:i:5", "te:i:11", "as:i:3", "xs:i:3",
            "si:i:0", "sc:Z:3M1D2M", "cl:i:2", "ci:i:1", "cn:i:2", "AS:i:-3", "NM:i:1", "SA:Z:chrB,6,+,2H3M1D2M3H,0,1;chrA,32,-,1H2M7H,0,0"),
        rec("r3", 2320, "chrA", 32, 0, "1H2M7H", "*", 0, 0, "GC", "87", "qs:i:7", "qe:i:9", "ts:i:7", "te:i:9", "as:i:3", "xs:i:3",
            "si:i:1", "sc:Z:2M", "cl:i:2", "ci:i:1", "cn:i:2", "AS:i:2", "NM:i:0", "SA:Z:chrB,6,+,2H3M1D2M3H,0,1;chrA,32,-,1H2M7H,0,0"),
    ]))

# ---- K4 ------------------------------------------------------------------------------------------------------------------
# The same chains with --soft-clip: hard_clip = false, so every record takes the (_, false) arms (mod.rs:781, 790-798):
# whole SEQ / QUAL, the forward CIGAR is NOT reversed, clips are S.
CASES.append(dict(
    name="K4_secondary_soft_clip",
    cites="mod.rs:690,781,790-798,822-846",
    targets=[["chrA", 40], ["chrB", 30]], double_strand=True, scoring=[1, -4, -6, -2],
    header="r3", bases="ACCGTTAGCA", quals="0123456789", pre_align_score=None, sam_opts=dict(soft_clip=True),
    chains=K3_CHAINS,
    expect=[
        rec("r3", 0, "chrA", 11, 60, "10M", "*", 0, 0, "ACCGTTAGCA", "0123456789", "qs:i:0", "qe:i:10", "ts:i:10", "te:i:20", "as:i:10", "xs:i:3",
            "si:i:0", "sc:Z:10M", "cl:i:1", "ci:i:0", "cn:i:2", "AS:i:10", "NM:i:0", "SA:Z:chrA,11,+,10M,60,0"),
        rec("r3", 256, "chrB", 6, 0, "2S2M1D3M3S", "*", 0, 0, "ACCGTTAGCA", "0123456789", "qs:i:2", "qe:i:7", "ts:i:5", "te:i:11", "as:i:3", "xs:i:3",
            "si:i:0", "sc:Z:2M1D3M", "cl:i:2", "ci:i:1", "cn:i:2", "AS:i:-3", "NM:i:1", "SA:Z:chrB,6,+,2S2M1D3M3S,0,1;chrA,32,-,1S2M7S,0,0"),
        rec("r3", 2320, "chrA", 32, 0, "1S2M7S", "*", 0, 0, "TGCTAACGGT", "9876543210", "qs:i:7", "qe:i:9", "ts:i:7", "te:i:9", "as:i:3", "xs:i:3",
            "si:i:1", "sc:Z:2M", "cl:i:2", "ci:i:1", "cn:i:2", "AS:i:2", "NM:i:0", "SA:Z:chrB,6,+,2S2M1D3M3S,0,1;chrA,32,-,1S2M7S,0,0"),
    ]))

# ---- K5 ------------------------------------------------------------------------------------------------------------------
# No chains: one unmapped record (mod.rs:634-668): flag 4, no reference, MAPQ 0, empty CIGAR, SEQ as given, `xs` only when
# a pre-alignment score is passed.  The read name is the first whitespace-separated word of the header (mod.rs:612-619).
CASES.append(dict(
    name="K5_unmapped_with_prealign_score", cites="mod.rs:612-619,634-668",
    targets=[["chrA", 40]], double_strand=False, scoring=[1, -4, -6, -2],
    header="readX some description", bases="ACGT", quals=None, pre_align_score=37, sam_opts={}, chains=[],
    expect=[rec("readX", 4, "*", 0, 0, "*", "*", 0, 0, "ACGT", "*", "xs:i:37")]))
CASES.append(dict(
    name="K5b_unmapped_plain", cites="mod.rs:634-668",
    targets=[["chrA", 40]], double_strand=False, scoring=[1, -4, -6, -2],
    header="readY", bases="ACGT", quals="IIII", pre_align_score=None, sam_opts={}, chains=[],
    expect=[rec("readY", 4, "*", 0, 0, "*", "*", 0, 0, "ACGT", "IIII")]))

# ---- K6 ------------------------------------------------------------------------------------------------------------------
# --pick-primary score (mod.rs:707-714): sub 1 (score 6) beats sub 0 (score 3) although sub 0 is longer.  --filter-secondary
# with 60 %: min_score = 6 * 60 / 100 = 3.6 (f32, mod.rs:725-727); sub 0 (3 < 3.6) is dropped, the primary index is re-mapped
# to the position the primary takes in the filtered list (mod.rs:730-745): 0.  `cl` is the FILTERED length (mod.rs:911-914)
# and `si` the index in the filtered list.  The pre-alignment score alone gives xs (mod.rs:678-687).
CASES.append(dict(
    name="K6_pick_primary_score_filter_secondary", cites="mod.rs:678-687,707-746,893-914",
    targets=[["chrA", 40]], double_strand=False, scoring=[1, -4, -6, -2],
    header="r6", bases="ACGTACGTACGTAC", quals="IIIIIIIIIIIIII", pre_align_score=12,
    sam_opts=dict(use_eq_and_x=True, pick_primary=1, filter_secondary=True, filter_secondary_pct=60.0),
    chains=[dict(score=-1, xstart=0, xend=26, ystart=0, yend=14, xlen=40, ylen=14, start_contig_idx=0, end_contig_idx=0, length=14,
                 ops=[[M, 4, 0], [X, 1, 0], [M, 3, 0], [XJ, 0, 20], [M, 6, 0]])],
    expect=[rec("r6", 0, "chrA", 21, 60, "8S6=", "*", 0, 0, "ACGTACGTACGTAC", "IIIIIIIIIIIIII", "qs:i:8", "qe:i:14", "ts:i:20", "te:i:26", "as:i:-1",
                "xs:i:12", "si:i:0", "sc:Z:6=", "cl:i:1", "ci:i:0", "cn:i:1", "AS:i:6", "NM:i:0", "SA:Z:chrA,21,+,8S6=,60,0")]))

# ---- K7 ------------------------------------------------------------------------------------------------------------------
# A chain re-joined by split_at_y (alignment.rs:207-360) carries a Yjump: it closes a sub-alignment, skips read bases and
# keeps the contig position (sub_alignment.rs:104-126).  Primary = sub 1 (longer); rotate_right(1) of two strings swaps them.
CASES.append(dict(
    name="K7_yjump", cites="sub_alignment.rs:104-126,169-241; mod.rs:956",
    targets=[["chrA", 40]], double_strand=False, scoring=[1, -4, -6, -2],
    header="r7", bases="ACGTACGTACGT", quals=None, pre_align_score=None, sam_opts={},
    chains=[dict(score=9, xstart=12, xend=21, ystart=0, yend=12, xlen=40, ylen=12, start_contig_idx=0, end_contig_idx=0, length=9,
                 ops=[[M, 4, 0], [YJ, 3, 0], [M, 5, 0]])],
    expect=[
        rec("r7", 2048, "chrA", 13, 60, "4M8S", "*", 0, 0, "ACGTACGTACGT", "*", "qs:i:0", "qe:i:4", "ts:i:12", "te:i:16", "as:i:9",
            "si:i:0", "sc:Z:4M", "cl:i:2", "ci:i:0", "cn:i:1", "AS:i:4", "NM:i:0", "SA:Z:chrA,17,+,7S5M,60,0;chrA,13,+,4M8S,60,0"),
        rec("r7", 0, "chrA", 17, 60, "7S5M", "*", 0, 0, "ACGTACGTACGT", "*", "qs:i:7", "qe:i:12", "ts:i:16", "te:i:21", "as:i:9",
            "si:i:1", "sc:Z:5M", "cl:i:2", "ci:i:0", "cn:i:1", "AS:i:5", "NM:i:0", "SA:Z:chrA,17,+,7S5M,60,0;chrA,13,+,4M8S,60,0"),
    ]))

# ---- K8 ------------------------------------------------------------------------------------------------------------------
# A chain that starts with an Xjump (cigar `..J4=`): the sub-alignment the Xjump closes consumes no read base and is dropped
# (sub_alignment.rs:198-202 "ignore alignments that do not consume target bases"); one record remains.
CASES.append(dict(
    name="K8_leading_xjump_dropped", cites="sub_alignment.rs:81-103,198-202",
    targets=[["chrA", 20], ["chrB", 20]], double_strand=False, scoring=[1, -4, -6, -2],
    header="r8", bases="ACGT", quals="FFFF", pre_align_score=None, sam_opts={},
    chains=[dict(score=-6, xstart=0, xend=9, ystart=0, yend=4, xlen=20, ylen=4, start_contig_idx=0, end_contig_idx=1, length=4,
                 ops=[[XJ, 1, 5], [M, 4, 0]])],
    expect=[rec("r8", 0, "chrB", 6, 60, "4M", "*", 0, 0, "ACGT", "FFFF", "qs:i:0", "qe:i:4", "ts:i:5", "te:i:9", "as:i:-6",
                "si:i:0", "sc:Z:4M", "cl:i:1", "ci:i:0", "cn:i:1", "AS:i:4", "NM:i:0", "SA:Z:chrB,6,+,4M,60,0")]))

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sam_kats.json")
    with open(out, "w") as f:
        json.dump(CASES, f, indent=1)
    print(f"{len(CASES)} cases, {sum(len(c['expect']) for c in CASES)} records -> {out}")
