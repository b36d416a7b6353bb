// capi_impl.hpp — the extern "C" surface of include/stitch_b200.h over host::Aligner.
//
// Included exactly once per shared library after defining
//   STITCH_API(name)            the exported symbol name (product: stitch_##name)
//   stitch::host::Backend *stitch_make_backend(stitch::host::Aligner &, int device)
// The product library (capi.cu) binds the CUDA backend.  The test-only emulator library binds a
// CPU emulation backend under a different prefix; it is never linked into the product.
#pragma once
#include <cstdio>
#include <string>

#include "host_common.hpp"
#include "host_sam.hpp"

namespace stitch { namespace host { Backend *stitch_make_backend(Aligner &al, int device); } }

struct stitch_ctx { stitch::host::Aligner al; };
struct stitch_results { stitch::host::Results res; };

static thread_local std::string g_create_error;

#define STITCH_GUARD_BEGIN try {
#define STITCH_GUARD_END(ctx)                                                         \
    } catch (const stitch::host::Error &e) {                                          \
        (ctx)->al.last_error = e.what(); return e.code;                               \
    } catch (const std::bad_alloc &) {                                                \
        (ctx)->al.last_error = "out of host memory"; return STITCH_ERR_NOMEM;         \
    } catch (const std::exception &e) {                                               \
        (ctx)->al.last_error = e.what(); return STITCH_ERR_INTERNAL;                  \
    }

extern "C" {

int STITCH_API(create)(const stitch_opts *opts, const stitch_contig *contigs, uint32_t n_contigs, int device,
                       stitch_ctx **out) {
    if (!opts || !contigs || !out) { g_create_error = "null argument"; return STITCH_ERR_INVALID; }
    stitch_ctx *ctx = nullptr;
    try {
        ctx = new stitch_ctx();
        ctx->al.init(*opts, contigs, n_contigs);
        ctx->al.backend.reset(stitch::host::stitch_make_backend(ctx->al, device));
        *out = ctx;
        return STITCH_OK;
    } catch (const stitch::host::Error &e) {
        g_create_error = e.what(); delete ctx; return e.code;
    } catch (const std::exception &e) {
        g_create_error = e.what(); delete ctx; return STITCH_ERR_INTERNAL;
    }
}

static int run_batch(stitch_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads,
                     const uint32_t *subset_words, uint32_t subset_stride, stitch_results **out, bool raw) {
    if (!ctx) return STITCH_ERR_INVALID;
    STITCH_GUARD_BEGIN
    if (!out || (n_reads && (!bases || !offsets))) throw stitch::host::Error(STITCH_ERR_INVALID, "null argument");
    for (uint32_t r = 0; r < n_reads; ++r)
        if (offsets[r + 1] < offsets[r]) throw stitch::host::Error(STITCH_ERR_INVALID, "offsets not monotone");
    std::unique_ptr<stitch_results> r(new stitch_results());
    ctx->al.backend->stats.reset();
    if (raw) ctx->al.custom_batch(bases, offsets, n_reads, subset_words, subset_stride, r->res);
    else ctx->al.align_batch(bases, offsets, n_reads, subset_words, subset_stride, r->res);
    *out = r.release();
    return STITCH_OK;
    STITCH_GUARD_END(ctx)
}

int STITCH_API(align_batch)(stitch_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads,
                            const uint32_t *subset_words, uint32_t subset_stride, stitch_results **out) {
    return run_batch(ctx, bases, offsets, n_reads, subset_words, subset_stride, out, false);
}

int STITCH_API(custom_batch)(stitch_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads,
                             const uint32_t *subset_words, uint32_t subset_stride, stitch_results **out) {
    return run_batch(ctx, bases, offsets, n_reads, subset_words, subset_stride, out, true);
}

int STITCH_API(prealign_batch)(stitch_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads,
                               uint32_t *subset_words, uint32_t subset_stride, int32_t *best_scores) {
    if (!ctx) return STITCH_ERR_INVALID;
    STITCH_GUARD_BEGIN
    if (n_reads && (!bases || !offsets || !subset_words)) throw stitch::host::Error(STITCH_ERR_INVALID, "null argument");
    if (!ctx->al.opts.pre_align) throw stitch::host::Error(STITCH_ERR_INVALID, "the context was created without pre_align");
    if ((uint64_t)subset_stride * 32 < ctx->al.contigs.n_strands) throw stitch::host::Error(STITCH_ERR_INVALID, "subset_stride too small");
    ctx->al.backend->stats.reset();
    ctx->al.prealign_batch(bases, offsets, n_reads, subset_words, subset_stride, best_scores);
    return STITCH_OK;
    STITCH_GUARD_END(ctx)
}

uint32_t STITCH_API(results_n_reads)(const stitch_results *r) { return (uint32_t)r->res.first.size(); }
void STITCH_API(results_read)(const stitch_results *r, uint32_t read, uint64_t *first, uint32_t *count) {
    *first = r->res.first[read]; *count = r->res.count[read];
}
const stitch_chain *STITCH_API(results_chains)(const stitch_results *r, uint64_t *n) {
    *n = r->res.chains.size(); return r->res.chains.data();
}
const stitch_op *STITCH_API(results_ops)(const stitch_results *r, uint64_t *n) {
    *n = r->res.ops.size(); return r->res.ops.data();
}
void STITCH_API(free_results)(stitch_results *r) { delete r; }
int STITCH_API(results_prealign)(const stitch_results *r, uint32_t read, int32_t *score) {
    if (!r || read >= r->res.has_pre.size() || !r->res.has_pre[read]) return 0;
    if (score) *score = r->res.pre_score[read];
    return 1;
}
int STITCH_API(results_from_chains)(const stitch_chain *chains, uint32_t n_chains, const stitch_op *ops, uint64_t n_ops, stitch_results **out) {
    if (!out || (n_chains && !chains) || (n_ops && !ops)) return STITCH_ERR_INVALID;
    for (uint32_t c = 0; c < n_chains; ++c)
        if (chains[c].ops_offset > n_ops || chains[c].n_ops > n_ops - chains[c].ops_offset) return STITCH_ERR_INVALID;
    try {
        std::unique_ptr<stitch_results> r(new stitch_results());
        r->res.begin_read();
        for (uint32_t c = 0; c < n_chains; ++c) {
            stitch_chain ch = chains[c];
            const uint64_t src = ch.ops_offset;
            ch.ops_offset = r->res.ops.size();
            r->res.ops.insert(r->res.ops.end(), ops + src, ops + src + ch.n_ops);
            r->res.chains.push_back(ch);
            r->res.count.back() += 1;
        }
        *out = r.release();
        return STITCH_OK;
    } catch (const std::exception &) { return STITCH_ERR_NOMEM; }
}

int STITCH_API(format_sam)(stitch_ctx *ctx, const stitch_results *res, uint32_t read, const char *read_header, const uint8_t *bases,
                           const uint8_t *quals, uint32_t n_bases, int has_pre_align_score, int32_t pre_align_score,
                           const stitch_sam_opts *sopts, char **out_text) {
    if (!ctx) return STITCH_ERR_INVALID;
    STITCH_GUARD_BEGIN
    if (!res || !read_header || (!bases && n_bases) || !out_text) throw stitch::host::Error(STITCH_ERR_INVALID, "null argument");
    if (read >= res->res.first.size()) throw stitch::host::Error(STITCH_ERR_INVALID, "read index out of range");
    stitch::host::SamOpts o;
    if (sopts) {
        o.soft_clip = sopts->soft_clip != 0; o.use_eq_and_x = sopts->use_eq_and_x != 0; o.filter_secondary = sopts->filter_secondary != 0;
        o.pick_primary = sopts->pick_primary; o.filter_secondary_pct = sopts->filter_secondary_pct;
    }
    const std::string text = stitch::host::format_sam(ctx->al.contigs, ctx->al.opts.sc, o, read_header, bases, quals, n_bases,
                                                      res->res.chains.data() + res->res.first[read], res->res.count[read],
                                                      res->res.ops.data(), has_pre_align_score != 0, pre_align_score);
    char *buf = static_cast<char *>(std::malloc(text.size() + 1));
    if (!buf) throw std::bad_alloc();
    std::memcpy(buf, text.c_str(), text.size() + 1);
    *out_text = buf;
    return STITCH_OK;
    STITCH_GUARD_END(ctx)
}
void STITCH_API(free_text)(char *t) { std::free(t); }

int STITCH_API(get_stats)(const stitch_ctx *ctx, stitch_stats *out) {
    if (!ctx || !out) return STITCH_ERR_INVALID;
    const stitch::host::BackendStats &s = ctx->al.backend->stats;
    out->cells = s.cells; out->fills = s.fills; out->kernel_launches = s.launches;
    out->fill_ms = s.fill_ms; out->traceback_ms = s.tb_ms; out->total_ms = s.total_ms;
    out->h2d_bytes = s.h2d; out->d2h_bytes = s.d2h; out->traceback_bytes = s.tb_bytes;
    out->packed_fill_ms = s.packed_ms; out->wide_fill_ms = s.wide_ms; out->redo_fill_ms = s.redo_ms; out->tail_fill_ms = s.tail_ms; out->packed_launches = s.packed_launches;
    out->packed_cells = s.packed_cells; out->redo_fills = s.refills;
    out->tile_columns = s.tile_columns; out->quiet_tile_columns = s.quiet_tile_columns;
    out->prealign_ms = s.pre_ms; out->prealign_reads = s.pre_reads;
    return STITCH_OK;
}

int STITCH_API(set_max_inflight)(stitch_ctx *ctx, uint32_t max_reads) {
    if (!ctx) return STITCH_ERR_INVALID;
    ctx->al.backend->set_max_inflight(max_reads);
    return STITCH_OK;
}

void STITCH_API(destroy)(stitch_ctx *ctx) { delete ctx; }

const char *STITCH_API(last_error)(const stitch_ctx *ctx) {
    return ctx ? ctx->al.last_error.c_str() : g_create_error.c_str();
}

uint32_t STITCH_API(abi_version)(void) { return 3; }

}  // extern "C"
