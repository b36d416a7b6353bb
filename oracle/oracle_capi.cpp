// TEST INFRASTRUCTURE ONLY — C entry points over the CPU oracle for ctypes (tests/, bench.py's
// cpu_baseline and --impl reference legs, __graft_entry__.smoke()).  Results use the same
// stitch_chain / stitch_op records as the product's C ABI (include/stitch_b200.h) so the parity
// tests compare them field by field.
#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../include/stitch_b200.h"
#include "stitch_oracle.hpp"

using namespace oracle;

struct oracle_result {
    std::vector<stitch_chain> chains;
    std::vector<stitch_op> ops;
    std::vector<uint64_t> first;     // per read
    std::vector<uint32_t> count;     // per read
    std::vector<std::string> cigars; // per chain (Alignment::cigar())
    uint64_t cells = 0, fills = 0;
    double seconds = 0;
};

static thread_local std::string g_err;

static void append_chain(oracle_result &r, const Alignment &a) {
    stitch_chain c{};
    c.score = a.score;
    c.xstart = (uint32_t)a.xstart; c.xend = (uint32_t)a.xend;
    c.ystart = (uint32_t)a.ystart; c.yend = (uint32_t)a.yend;
    c.xlen = (uint32_t)a.xlen; c.ylen = (uint32_t)a.ylen;
    c.start_contig_idx = (uint32_t)a.start_contig_idx; c.end_contig_idx = (uint32_t)a.end_contig_idx;
    c.length = (uint32_t)a.length;
    c.ops_offset = r.ops.size();
    for (const Op &op : a.ops) {
        if (op.kind <= INS && !r.ops.empty() && r.ops.size() > c.ops_offset && r.ops.back().kind == op.kind) {
            r.ops.back().a += 1;
        } else if (op.kind <= INS) {
            r.ops.push_back(stitch_op{op.kind, 1, 0});
        } else {
            r.ops.push_back(stitch_op{op.kind, op.a, op.b});
        }
    }
    c.n_ops = (uint32_t)(r.ops.size() - c.ops_offset);
    r.chains.push_back(c);
    r.cigars.push_back(a.cigar());
}

static Alignment from_chain(const stitch_chain *c, const stitch_op *ops, int mode) {
    Alignment a;
    a.score = c->score; a.xstart = c->xstart; a.xend = c->xend; a.ystart = c->ystart; a.yend = c->yend;
    a.xlen = c->xlen; a.ylen = c->ylen; a.start_contig_idx = c->start_contig_idx;
    a.end_contig_idx = c->end_contig_idx; a.length = c->length; a.mode = (Mode)mode;
    for (uint32_t k = 0; k < c->n_ops; ++k) {
        const stitch_op &o = ops[k];
        if (o.kind <= INS) for (uint32_t t = 0; t < o.a; ++t) a.ops.push_back(Op{(uint8_t)o.kind});
        else a.ops.push_back(Op{(uint8_t)o.kind, o.a, o.b});
    }
    return a;
}

static Options to_options(const stitch_opts *o) {
    Options p;
    p.mode = (Mode)o->mode;
    p.match_score = o->match_score; p.mismatch_score = o->mismatch_score;
    p.gap_open = o->gap_open; p.gap_extend = o->gap_extend;
    p.jump_same = o->jump_same; p.jump_opp = o->jump_opp; p.jump_inter = o->jump_inter;
    p.double_strand = o->double_strand; p.circular = o->circular; p.suboptimal = o->suboptimal;
    p.circular_slop = o->circular_slop; p.suboptimal_pct = o->suboptimal_pct;
    return p;
}

struct oracle_aligner {
    Options opts;
    std::vector<std::string> names;
    std::vector<std::vector<uint8_t>> fwd;
    std::unique_ptr<Aligners> al;
};

extern "C" {

const char *oracle_last_error(void) { return g_err.c_str(); }
// 0 (default): the reference's traceback layout (what the CPU baseline is timed on); 1: column-major (parity checks only)
void oracle_set_checker_layout(int colmajor) { SingleContig::g_checker_layout = colmajor != 0; }

// scoring = {match, mismatch, gap_open, gap_extend, jump}; mode 0..3 = the SCA mode wrappers
// (single_contig_aligner.rs:733-872), which also filter the free clip ops.
int oracle_sca(int mode, const int32_t *scoring, int circular, const uint8_t *x, int64_t m,
               const uint8_t *y, int64_t n, oracle_result **out) {
    try {
        SingleContig sc;
        sc.sc.match = scoring[0]; sc.sc.mismatch = scoring[1];
        sc.sc.gap_open = scoring[2]; sc.sc.gap_extend = scoring[3];
        sc.sc.jump_same = sc.sc.jump_opp = sc.sc.jump_inter = scoring[4];
        sc.circular = circular != 0;
        Alignment a = sc.with_mode((Mode)mode, x, m, y, n);
        auto r = new oracle_result();
        r->first.push_back(0); r->count.push_back(1);
        append_chain(*r, a);
        r->cells = (uint64_t)m * (uint64_t)n; r->fills = 1;
        *out = r;
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// MultiContigAligner::custom over explicitly listed contig-strands.
// scoring (per call) = {match, mismatch, gap_open, gap_extend, jump_same, jump_opp, jump_inter,
//                       xclip_prefix, xclip_suffix, yclip_prefix, yclip_suffix}
int oracle_mca(uint32_t n_contigs, const uint8_t *const *seqs, const int64_t *lens, const char *const *names,
               const uint8_t *is_forward, const uint8_t *circular, const int32_t *scoring,
               const uint8_t *y, int64_t n, const uint8_t *subset, oracle_result **out) {
    try {
        Scoring s;
        s.match = scoring[0]; s.mismatch = scoring[1]; s.gap_open = scoring[2]; s.gap_extend = scoring[3];
        s.jump_same = scoring[4]; s.jump_opp = scoring[5]; s.jump_inter = scoring[6];
        s.xclip_prefix = scoring[7]; s.xclip_suffix = scoring[8]; s.yclip_prefix = scoring[9]; s.yclip_suffix = scoring[10];
        MultiContig mc;
        for (uint32_t c = 0; c < n_contigs; ++c)
            mc.add_contig(names[c], is_forward[c] != 0, seqs[c], lens[c], circular[c] != 0, s);
        std::vector<bool> sub;
        if (subset) sub.assign(subset, subset + n_contigs);
        Alignment a = mc.custom_with_subset(y, n, subset ? &sub : nullptr);
        auto r = new oracle_result();
        r->first.push_back(0); r->count.push_back(1);
        append_chain(*r, a);
        r->cells = mc.cells_filled; r->fills = mc.fills;
        *out = r;
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

int oracle_split_at_y(const stitch_chain *chain, const stitch_op *ops, int mode, int64_t y_pivot, oracle_result **out) {
    try {
        Alignment a = from_chain(chain, ops, mode).split_at_y(y_pivot);
        auto r = new oracle_result();
        r->first.push_back(0); r->count.push_back(1);
        append_chain(*r, a);
        *out = r;
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// Alignment::cigar() of an arbitrary chain (used to render product results in the tests).
int oracle_cigar(const stitch_chain *chain, const stitch_op *ops, char *buf, size_t buf_len) {
    try {
        std::string s = from_chain(chain, ops, CUSTOM).cigar();
        if (s.size() + 1 > buf_len) return -2;
        std::memcpy(buf, s.c_str(), s.size() + 1);
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

int oracle_aligner_create(const stitch_opts *opts, const stitch_contig *contigs, uint32_t n_contigs, oracle_aligner **out) {
    try {
        auto h = new oracle_aligner();
        h->opts = to_options(opts);
        for (uint32_t c = 0; c < n_contigs; ++c) {
            h->names.emplace_back(contigs[c].name);
            h->fwd.emplace_back(contigs[c].fwd, contigs[c].fwd + contigs[c].len);
        }
        h->al.reset(new Aligners(h->opts, h->names, h->fwd));
        *out = h;
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}
void oracle_aligner_destroy(oracle_aligner *h) { delete h; }

// raw != 0: MultiContigAligner::custom_with_subset (one chain, clips kept); else Aligners::align.
// n_threads > 1: one independent Aligners per thread (fg-stitch-cli/src/commands/align.rs:345-379).
int oracle_aligner_batch(oracle_aligner *h, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads,
                         const uint32_t *subset_words, uint32_t subset_stride, int raw, int n_threads,
                         oracle_result **out) {
    try {
        const size_t C = h->al->mc.contigs.size();
        std::vector<std::vector<Alignment>> per_read(n_reads);
        std::vector<std::string> errs((size_t)std::max(1, n_threads));
        std::atomic<uint32_t> next{0};
        std::atomic<uint64_t> cells{0}, fills{0};
        auto t0 = std::chrono::steady_clock::now();
        auto work = [&](int tid, Aligners *al) {
            try {
                for (;;) {
                    uint32_t r = next.fetch_add(1);
                    if (r >= n_reads) break;
                    std::vector<bool> sub;
                    const std::vector<bool> *subp = nullptr;
                    if (subset_words) {
                        sub.assign(C, false);
                        bool any = false;
                        for (size_t c = 0; c < C; ++c) {
                            bool b = (subset_words[(size_t)r * subset_stride + c / 32] >> (c % 32)) & 1u;
                            sub[c] = b; any = any || b;
                        }
                        if (any) subp = &sub;
                    }
                    const uint8_t *q = bases + offsets[r];
                    int64_t n = (int64_t)(offsets[r + 1] - offsets[r]);
                    if (raw) {
                        std::vector<uint8_t> up(q, q + n);
                        for (auto &b : up) if (b >= 'a' && b <= 'z') b = (uint8_t)(b - 32);
                        per_read[r].push_back(al->mc.custom_with_subset(up.data(), n, subp));
                    } else {
                        per_read[r] = al->align(q, n, subp);
                    }
                }
                cells += al->mc.cells_filled; fills += al->mc.fills;
                al->mc.cells_filled = 0; al->mc.fills = 0;
            } catch (const std::exception &e) { errs[(size_t)tid] = e.what(); }
        };
        if (n_threads <= 1) {
            work(0, h->al.get());
        } else {
            std::vector<std::unique_ptr<Aligners>> als;
            for (int t = 0; t < n_threads; ++t) als.emplace_back(new Aligners(h->opts, h->names, h->fwd));
            t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t, als[(size_t)t].get());
            for (auto &t : th) t.join();
        }
        auto t1 = std::chrono::steady_clock::now();
        for (auto &e : errs) if (!e.empty()) { g_err = e; return -1; }
        auto r = new oracle_result();
        for (uint32_t k = 0; k < n_reads; ++k) {
            r->first.push_back(r->chains.size());
            r->count.push_back((uint32_t)per_read[k].size());
            for (auto &a : per_read[k]) append_chain(*r, a);
        }
        r->cells = cells; r->fills = fills;
        r->seconds = std::chrono::duration<double>(t1 - t0).count();
        *out = r;
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

uint32_t oracle_result_n_reads(const oracle_result *r) { return (uint32_t)r->first.size(); }
void oracle_result_read(const oracle_result *r, uint32_t read, uint64_t *first, uint32_t *count) {
    *first = r->first[read]; *count = r->count[read];
}
const stitch_chain *oracle_result_chains(const oracle_result *r, uint64_t *n) { *n = r->chains.size(); return r->chains.data(); }
const stitch_op *oracle_result_ops(const oracle_result *r, uint64_t *n) { *n = r->ops.size(); return r->ops.data(); }
const char *oracle_result_cigar(const oracle_result *r, uint64_t chain) { return r->cigars[chain].c_str(); }
uint64_t oracle_result_cells(const oracle_result *r) { return r->cells; }
uint64_t oracle_result_fills(const oracle_result *r) { return r->fills; }
double oracle_result_seconds(const oracle_result *r) { return r->seconds; }
void oracle_result_free(oracle_result *r) { delete r; }

}  // extern "C"
