// host_sam.hpp — SAM records of one read from its chains: the reference's SubAlignmentBuilder
// (fg-stitch-lib/src/align/sub_alignment.rs:36-241) and SamRecordFormatter::format
// (fg-stitch-lib/src/align/aligners/mod.rs:606-973), restated on the run-length encoded operations of
// stitch_results.  Host code of the "next" row (f)-1 of SURVEY.md section 8: it turns what the GPU path returns
// into the `stitch align` output records (flags, positions, CIGARs, the custom chain tags qs/qe/ts/te/as/xs/
// si/sc/cl/ci/cn, AS, NM, SA), rendered as SAM text lines.  noodles' byte-level encoding (BAM, tag typing of
// the binary form, its reading of FASTQ quality bytes) is not restated: qualities are passed through as given.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "host_common.hpp"

namespace stitch {
namespace host {

struct SamOpts {
    bool soft_clip = false, use_eq_and_x = false, filter_secondary = false;
    int pick_primary = 0;                 // 0 = query-length (default), 1 = score   (align/mod.rs:15-19)
    float filter_secondary_pct = 10.0f;
};

struct SubAln {                           // sub_alignment.rs:10-19, after the x/y swap of build(.., swap = true)
    size_t contig_idx = 0, query_start = 0, query_end = 0, target_start = 0, target_end = 0;
    std::vector<std::pair<char, size_t>> cigar;
    int32_t score = 0, num_edits = 0;
};

inline std::string cigar_string(const std::vector<std::pair<char, size_t>> &c) {
    std::string s;
    for (const auto &e : c) { s += std::to_string(e.second); s += e.first; }
    return s;
}

// SubAlignmentBuilder::build(chain, swap = true, scoring).  The builder walks unit operations; its "query" is
// the contig axis (x) and its "target" the read axis (y) until the final swap (sub_alignment.rs:214-240).
inline std::vector<SubAln> build_subs(const stitch_chain &ch, const stitch_op *ops, bool use_eq_and_x, const Scoring &sc) {
    if (ch.n_ops == 0) throw Error(STITCH_ERR_INTERNAL, "chain without operations (the reference indexes operations[0])");
    struct U { uint32_t kind, a, b; };
    auto same = [&](const U &l, const U &c) {
        if (l.kind == c.kind && (l.kind <= OP_INS || (l.a == c.a && l.b == c.b))) return true;
        if (!use_eq_and_x && ((l.kind == OP_SUBST && c.kind == OP_MATCH) || (l.kind == OP_MATCH && c.kind == OP_SUBST))) return true;
        return false;
    };
    const char match_kind = use_eq_and_x ? '=' : 'M', mismatch_kind = use_eq_and_x ? 'X' : 'M';
    std::vector<std::pair<char, size_t>> elements;
    size_t q_start = ch.xstart, t_start = ch.ystart, q_off = ch.xstart, t_off = ch.ystart, contig = ch.start_contig_idx;
    int32_t score = 0, num_edits = 0;
    std::vector<SubAln> out;
    auto snapshot = [&]() {
        SubAln s; s.contig_idx = contig; s.query_start = q_start; s.query_end = q_off; s.target_start = t_start; s.target_end = t_off;
        s.cigar = elements; s.score = score; s.num_edits = num_edits;
        return s;
    };
    // add_op (sub_alignment.rs:49-132): returns true when a sub-alignment was emitted into `emitted`
    auto add_op = [&](const U &op, size_t len, SubAln &emitted) -> bool {
        switch (op.kind) {
        case OP_MATCH: score += sc.match * (int32_t)len; q_off += len; t_off += len; elements.emplace_back(match_kind, len); return false;
        case OP_SUBST: score += sc.mismatch * (int32_t)len; q_off += len; t_off += len; elements.emplace_back(mismatch_kind, len); return false;
        case OP_DEL: score += sc.o + sc.e * (int32_t)len; t_off += len; elements.emplace_back('D', len); return false;
        case OP_INS: score += sc.o + sc.e * (int32_t)len; q_off += len; elements.emplace_back('I', len); return false;
        case OP_XJUMP:
            emitted = snapshot();
            elements.clear(); contig = op.a; t_start = t_off; q_start = op.b; q_off = op.b; score = 0; num_edits = 0;
            return true;
        case OP_YJUMP:
            emitted = snapshot();
            elements.clear(); t_off += op.a; t_start = t_off; q_start = q_off; score = 0; num_edits = 0;
            return true;
        default:
            if (len != 1) throw Error(STITCH_ERR_INTERNAL, "repeated clip operation");   // assert!(op_len == 1)
            return false;
        }
    };
    // unit operations, expanded lazily from the runs
    U last{ops[0].kind, ops[0].kind <= OP_INS ? 0u : ops[0].a, ops[0].kind <= OP_INS ? 0u : ops[0].b};
    size_t op_len = 0;
    for (uint32_t r = 0; r < ch.n_ops; ++r) {
        const bool base = ops[r].kind <= OP_INS;
        const U u{ops[r].kind, base ? 0u : ops[r].a, base ? 0u : ops[r].b};
        const uint32_t reps = base ? ops[r].a : 1u;
        for (uint32_t t = 0; t < reps; ++t) {
            if (u.kind == OP_SUBST || u.kind == OP_INS || u.kind == OP_DEL) num_edits += 1;   // before the flush below (sub_alignment.rs:187-193)
            if (same(last, u)) op_len += 1;
            else {
                SubAln e;
                if (add_op(last, op_len, e) && e.target_start < e.target_end) out.push_back(std::move(e));   // must consume read bases
                op_len = 1;
            }
            last = u;
        }
    }
    {
        SubAln e;
        if (add_op(last, op_len, e)) out.push_back(std::move(e));
        else out.push_back(snapshot());
    }
    for (SubAln &s : out) {   // swap: query <-> target, I <-> D (sub_alignment.rs:157-167, 226-237)
        std::swap(s.query_start, s.target_start); std::swap(s.query_end, s.target_end);
        for (auto &e : s.cigar) e.first = e.first == 'D' ? 'I' : (e.first == 'I' ? 'D' : e.first);
    }
    return out;
}

inline std::string read_name_of(const std::string &header) {   // header_to_name, mod.rs:612-619
    size_t b = 0;
    while (b < header.size() && std::isspace((unsigned char)header[b])) ++b;
    size_t e = b;
    while (e < header.size() && !std::isspace((unsigned char)header[e])) ++e;
    if (e == b) throw Error(STITCH_ERR_INVALID, "empty read name");
    return header.substr(b, e - b);
}

// SamRecordFormatter::format (mod.rs:622-972): SAM text lines (no trailing newline on the last one).
inline std::string format_sam(const Contigs &contigs, const Scoring &sc, const SamOpts &o, const std::string &header,
                              const uint8_t *bases, const uint8_t *quals, size_t n, const stitch_chain *chains, size_t n_chains,
                              const stitch_op *all_ops, bool has_pre, int32_t pre_score) {
    const std::string name = read_name_of(header);
    const size_t T = contigs.n_targets;
    auto seq_of = [&](size_t lo, size_t hi, bool rc) {
        std::vector<uint8_t> v(bases + lo, bases + hi);
        if (rc) v = revcomp(v);
        return std::string(v.begin(), v.end());
    };
    auto qual_of = [&](size_t lo, size_t hi, bool rev) {
        if (!quals) return std::string("*");
        std::string q(reinterpret_cast<const char *>(quals) + lo, reinterpret_cast<const char *>(quals) + hi);
        if (rev) std::reverse(q.begin(), q.end());
        return q.empty() ? std::string("*") : q;
    };
    std::string out;
    if (n_chains == 0) {   // unmapped record (mod.rs:634-668)
        out = name + "\t4\t*\t0\t0\t*\t*\t0\t0\t" + (n ? seq_of(0, n, false) : std::string("*")) + "\t" + qual_of(0, n, false);
        if (has_pre) out += "\txs:i:" + std::to_string(pre_score);
        return out;
    }
    int32_t primary_alignment_score = MIN_SCORE;
    bool has_sub = false; int32_t suboptimal = 0;
    for (size_t c = 1; c < n_chains; ++c) { suboptimal = has_sub ? std::max(suboptimal, chains[c].score) : chains[c].score; has_sub = true; }
    if (has_pre) { suboptimal = has_sub ? std::max(suboptimal, pre_score) : pre_score; has_sub = true; }

    for (size_t chain_idx = 0; chain_idx < n_chains; ++chain_idx) {
        const stitch_chain &ch = chains[chain_idx];
        const bool hard_clip = !o.soft_clip;
        std::vector<SubAln> subs = build_subs(ch, all_ops + ch.ops_offset, o.use_eq_and_x, sc);
        // the sub-alignment without the supplementary flag: max_by_key keeps the LAST maximum (mod.rs:699-714)
        size_t primary = 0;
        for (size_t k = 1; k < subs.size(); ++k) {
            const auto lenk = (int64_t)subs[k].query_end - (int64_t)subs[k].query_start, lenp = (int64_t)subs[primary].query_end - (int64_t)subs[primary].query_start;
            const bool ge = o.pick_primary == 0 ? (lenk > lenp || (lenk == lenp && subs[k].score >= subs[primary].score))
                                                : (subs[k].score > subs[primary].score || (subs[k].score == subs[primary].score && lenk >= lenp));
            if (ge) primary = k;
        }
        if (chain_idx == 0) primary_alignment_score = subs[primary].score;
        if (o.filter_secondary) {   // mod.rs:723-744
            const float min_score = (float)primary_alignment_score * o.filter_secondary_pct / 100.0f;
            std::vector<SubAln> kept;
            for (size_t old = 0; old < subs.size(); ++old) {
                if (old == primary) primary = kept.size();
                if ((float)subs[old].score >= min_score) kept.push_back(subs[old]);
            }
            subs.swap(kept);
        }
        std::vector<std::string> lines, sa;
        for (size_t sub_idx = 0; sub_idx < subs.size(); ++sub_idx) {
            const SubAln &s = subs[sub_idx];
            const bool is_supp = sub_idx != primary, is_sec = chain_idx > 0;
            if (s.contig_idx >= 2 * T) throw Error(STITCH_ERR_INTERNAL, "sub-alignment contig index out of range");
            const bool fwd = s.contig_idx < T;
            int flags = 0;
            if (!fwd) flags |= 0x10;
            if (is_sec) flags |= 0x100;
            if (is_supp) flags |= 0x800;
            const bool clip_seq = hard_clip && is_sec;
            std::vector<std::pair<char, size_t>> cig = s.cigar;
            if (!(fwd && !clip_seq)) std::reverse(cig.begin(), cig.end());   // also the (forward, hard-clipped) case: mod.rs:782-789
            const std::string sub_cigar = cigar_string(cig);
            const size_t lo = clip_seq ? s.query_start : 0, hi = clip_seq ? s.query_end : n;
            const std::string seq = seq_of(lo, hi, !fwd), qual = qual_of(lo, hi, !fwd);
            const char clip_op = clip_seq ? 'H' : 'S';
            std::vector<std::pair<char, size_t>> full;
            const size_t pre = fwd ? s.query_start : n - s.query_end, suf = fwd ? n - s.query_end : s.query_start;
            if (pre > 0) full.emplace_back(clip_op, pre);
            full.insert(full.end(), cig.begin(), cig.end());
            if (suf > 0) full.emplace_back(clip_op, suf);
            const std::string full_cigar = cigar_string(full);
            const size_t ref_id = s.contig_idx % T;
            const size_t ref_start = fwd ? s.target_start + 1 : contigs.len[ref_id] - s.target_end + 1;
            const int mapq = chain_idx == 0 ? 60 : 0;
            std::string l = name + "\t" + std::to_string(flags) + "\t" + contigs.names[ref_id] + "\t" + std::to_string(ref_start) + "\t" +
                            std::to_string(mapq) + "\t" + (full_cigar.empty() ? "*" : full_cigar) + "\t*\t0\t0\t" + (seq.empty() ? "*" : seq) + "\t" + qual;
            l += "\tqs:i:" + std::to_string(s.query_start) + "\tqe:i:" + std::to_string(s.query_end) + "\tts:i:" + std::to_string(s.target_start) +
                 "\tte:i:" + std::to_string(s.target_end) + "\tas:i:" + std::to_string(ch.score);
            if (has_sub) l += "\txs:i:" + std::to_string(suboptimal);
            l += "\tsi:i:" + std::to_string(sub_idx) + "\tsc:Z:" + sub_cigar + "\tcl:i:" + std::to_string(subs.size()) + "\tci:i:" +
                 std::to_string(chain_idx) + "\tcn:i:" + std::to_string(n_chains) + "\tAS:i:" + std::to_string(s.score) + "\tNM:i:" +
                 std::to_string(s.num_edits);
            lines.push_back(std::move(l));
            sa.push_back(contigs.names[ref_id] + "," + std::to_string(ref_start) + "," + (fwd ? "+" : "-") + "," + full_cigar + "," +
                         std::to_string(mapq) + "," + std::to_string(s.num_edits));
        }
        if (!sa.empty()) std::rotate(sa.rbegin(), sa.rbegin() + (primary % sa.size()), sa.rend());   // rotate_right(primary): mod.rs:956
        std::string sa_all;
        for (size_t k = 0; k < sa.size(); ++k) { if (k) sa_all += ";"; sa_all += sa[k]; }
        for (std::string &l : lines) {
            if (!out.empty()) out += "\n";
            out += l + "\tSA:Z:" + sa_all;
        }
    }
    return out;
}

}  // namespace host
}  // namespace stitch
