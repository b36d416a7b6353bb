// TEST INFRASTRUCTURE ONLY — CPU emulation of the CUDA kernels' tile / lane / round structure
// over the product's own __host__ __device__ DP core (stitch_b200/csrc/dp_core.h) and host driver
// (host_common.hpp).  It lets the CPU-only test tier fuzz the decomposition (pass A, insertion-chain
// scan, pass B, row-m finalize, checkpoints, windowed y-suffix tracking, fix-up, unit re-fill +
// resumable walk, re-alignment driver) against the oracle.  Exported under the emul_ prefix; the
// product library never links or loads this.
#define STITCH_API(name) emul_##name
#include "../../stitch_b200/csrc/capi_impl.hpp"

#include <cstdlib>
#include <vector>

namespace stitch {
namespace host {

#ifndef EMUL_WARPS
#define EMUL_WARPS 4
#endif

static uint32_t env_u32(const char *name, uint32_t dflt) {
    const char *v = std::getenv(name);
    return v ? (uint32_t)std::strtoul(v, nullptr, 10) : dflt;
}

// One column of a set of contigs (all contigs of the layout in the fill; one contig in a unit
// re-fill).  Mirrors fill_kernel's per-column body.
struct ColumnArgs {
    const Scoring *sc;
    const ContigEntry *ent; uint32_t C; uint32_t n_tiles;   // tile_start relative to the state arrays
    const uint8_t *bases; const uint8_t *read;
    uint32_t j, n;
    const CellState *prev; CellState *curr; CellState *ck_state;
    uint8_t *tb_col;
    ColRec *colrec_col;          // nullable
    const JumpInfo *J;           // per contig
    int32_t *Sm; uint32_t *slm, *tbm;                 // row-m summaries (in: column j-1, out: column j)
    int32_t *cm; uint32_t *cml, *cmk;                 // column best (out), nullable
    bool track; SnRec *sn; bool lastcol; LastCell *last;
};

static void emul_column(const ColumnArgs &A) {
    const Scoring &sc = *A.sc;
    const uint32_t C = A.C, j = A.j, n = A.n;
    const int W = EMUL_WARPS;
    const Row0 r0 = row0_at(sc, j, n), r0p = row0_at(sc, j - 1, n);
    ColConst cc; cc.j = j; cc.n = n; cc.q = A.read[j - 1];
    cc.xclip_score = sc.xp + std::max(sc.yp, sc.o + sc.e * (int32_t)j);
    cc.sl0j = r0.sl;
    std::vector<LaneA> la((size_t)W * 32);
    std::vector<ICarry> excl((size_t)W * 32), tileagg((size_t)W), tcarry((size_t)W);
    std::vector<XsPart> xs(C); std::vector<CmPart> cmp(C); std::vector<RowM> rowm(C);
    for (uint32_t a = 0; a < C; ++a) { xs_init(xs[a]); cm_init(cmp[a]); }
    ICarry round_carry{MIN_SCORE, 0, 0};
    for (uint32_t t0 = 0; t0 < A.n_tiles; t0 += (uint32_t)W) {
        const int nw = (int)std::min<uint32_t>((uint32_t)W, A.n_tiles - t0);
        std::vector<TileCtx> tcs((size_t)nw);
        for (int w = 0; w < nw; ++w) {   // phase A + warp scan
            const uint32_t tile = t0 + (uint32_t)w;
            uint32_t a = 0;
            while (!(tile >= A.ent[a].tile_start && tile < A.ent[a].tile_start + A.ent[a].ntiles)) ++a;
            const ContigEntry &en = A.ent[a];
            TileCtx tc; tc.a = a; tc.self_idx = en.contig_idx; tc.m = en.m; tc.tile = tile;
            tc.tile_in_contig = tile - en.tile_start; tc.J = A.J[a]; tc.circular = en.circular != 0;
            tc.wrap_src_ok = A.tbm[a] != TB_XCLIP_SUFFIX; tc.Sm_prev = A.Sm[a]; tc.slm_prev = A.slm[a];
            tcs[(size_t)w] = tc;
            for (uint32_t lane = 0; lane < 32; ++lane) {
                const uint32_t row0 = tc.tile_in_contig * TILE + lane * STRIP + 1;
                LaneA &out = la[(size_t)w * 32 + lane];
                out.has_m = 0; out.agg = ICarry{MIN_SCORE, 0, 0};
                if (row0 > en.m) continue;
                CellState up[STRIP]; uint8_t x[STRIP];
                for (int k = 0; k < STRIP; ++k) {
                    up[k] = A.prev[state_index(tile, lane, (uint32_t)k)];
                    const uint32_t i = row0 + (uint32_t)k;
                    x[k] = i <= en.m ? A.bases[en.seq_off + i - 1] : 0;
                }
                int32_t dgS; uint32_t dgsl;
                if (row0 == 1) { dgS = r0p.S; dgsl = r0p.sl; }
                else { const CellState &p = A.prev[row_index(en, row0 - 1)]; dgS = p.S; dgsl = p.sl; }
                lane_pass_a(sc, cc, tc, row0, up, dgS, dgsl, x, out, &rowm[a]);
            }
            ICarry cur[32];
            for (int l = 0; l < 32; ++l) cur[l] = la[(size_t)w * 32 + (size_t)l].agg;
            for (int d = 1; d < 32; d <<= 1) {
                ICarry nxt[32];
                for (int l = 0; l < 32; ++l)
                    nxt[l] = l >= d ? icarry_combine(cur[l - d], (uint32_t)(d * STRIP), sc.e, cur[l]) : cur[l];
                for (int l = 0; l < 32; ++l) cur[l] = nxt[l];
            }
            for (int l = 0; l < 32; ++l) excl[(size_t)w * 32 + (size_t)l] = l ? cur[l - 1] : ICarry{MIN_SCORE, 0, 0};
            tileagg[(size_t)w] = cur[31];
        }
        for (int w = 0; w < nw; ++w) {   // carry folding within the round
            const TileCtx &tc = tcs[(size_t)w];
            int w0 = w;
            while (w0 > 0 && tcs[(size_t)(w0 - 1)].a == tc.a) --w0;
            ICarry c = tcs[(size_t)w0].tile_in_contig == 0 ? icarry_row1(sc, r0) : round_carry;
            for (int u = w0; u < w; ++u) c = icarry_combine(c, TILE, sc.e, tileagg[(size_t)u]);
            tcarry[(size_t)w] = c;
        }
        round_carry = icarry_combine(tcarry[(size_t)(nw - 1)], TILE, sc.e, tileagg[(size_t)(nw - 1)]);
        for (int w = 0; w < nw; ++w) {   // phase B
            const TileCtx &tc = tcs[(size_t)w];
            const ContigEntry &en = A.ent[tc.a];
            for (uint32_t lane = 0; lane < 32; ++lane) {
                const uint32_t row0 = tc.tile_in_contig * TILE + lane * STRIP + 1;
                if (row0 > en.m) continue;
                LaneA &a_ = la[(size_t)w * 32 + lane];
                ICarry cin = lane == 0 ? tcarry[(size_t)w]
                                       : icarry_combine(tcarry[(size_t)w], lane * STRIP, sc.e, excl[(size_t)w * 32 + lane]);
                uint8_t x[STRIP];
                for (int k = 0; k < STRIP; ++k) {
                    const uint32_t i = row0 + (uint32_t)k;
                    x[k] = i <= en.m ? A.bases[en.seq_off + i - 1] : 0;
                }
                LaneB lb;
                lane_pass_b(sc, cc, tc, row0, lane, a_, cin, A.curr, A.ck_state, A.tb_col, A.track, A.sn, A.lastcol, A.last, x,
                            lb, &rowm[tc.a]);
                xs[tc.a] = xs_merge(xs[tc.a], lb.xs);
                cmp[tc.a] = cm_merge(cmp[tc.a], lb.cm);
            }
        }
    }
    for (uint32_t a = 0; a < C; ++a) {
        ContigColOut o = contig_finalize(sc, cc, A.ent[a], a, C, rowm[a], xs[a], cmp[a], r0, A.J[a], A.curr, A.ck_state, A.tb_col,
                                         A.colrec_col, A.track, A.sn, A.lastcol, A.last);
        if (A.cm) { A.cm[a] = o.cm.S; A.cmk[a] = o.cm.row; A.cml[a] = o.cm.sl; }
        A.Sm[a] = o.Sm; A.slm[a] = o.slm; A.tbm[a] = o.s_tb_m;
    }
}

struct EmulBackend : Backend {
    Aligner &al;
    uint32_t dump_seq = 0;
    uint32_t K, WINDOW;
    explicit EmulBackend(Aligner &a) : al(a) {
        K = std::max<uint32_t>(1, env_u32("EMUL_K", 7));          // checkpoint spacing (columns)
        WINDOW = std::max<uint32_t>(1, env_u32("EMUL_WINDOW", 6));  // y-suffix trackers kept for the last WINDOW columns
    }

    struct Fill {
        std::vector<ColRec> colrec; std::vector<LastCell> last; std::vector<SnRec> sn;
        std::vector<CellState> ck_state; std::vector<CkSum> ck_sum; std::vector<int32_t> gcol;
        bool need_full_track = false;
    };

    void fill(const Job &job, const Layout &L, uint32_t track_from, Fill &F) {
        const Scoring &sc = al.opts.sc;
        const uint32_t C = (uint32_t)L.ent.size(), PM = L.PM(), n = job.n;
        const uint8_t *bases = al.contigs.blob.data();
        const uint32_t nb = (n + K - 1) / K;
        std::vector<CellState> st[2];
        st[0].assign(PM, CellState{MIN_SCORE, MIN_SCORE, 0, 0}); st[1] = st[0];
        F.colrec.assign((size_t)(n + 1) * C, ColRec{});
        F.last.assign(PM, LastCell{});
        F.sn.assign(PM, SnRec{0x3fffffff, 7, 7, 7});   // garbage unless initialised below
        F.ck_state.assign((size_t)(nb ? nb - 1 : 0) * PM, CellState{});
        F.ck_sum.assign((size_t)(nb ? nb - 1 : 0) * C, CkSum{});
        F.gcol.assign((size_t)n + 1, MIN_SCORE);
        stats.cells += (uint64_t)L.cells_per_col * n;
        stats.fills += 1;
        const bool any_track = track_from <= n;
        std::vector<int32_t> cm(C), Sm(C);
        std::vector<uint32_t> cml(C), cmk(C), slm(C), tbm(C);
        for (uint32_t a = 0; a < C; ++a) {
            const ContigEntry &en = L.ent[a];
            CmPart part; cm_init(part); cm_add(part, 0, 0, 0);
            for (uint32_t i = 1; i <= en.m; ++i) {
                Col0 c0 = col0_at(sc, i, en.m);
                st[0][row_index(en, i)] = CellState{c0.S, MIN_SCORE, c0.sl, 0};
                if (any_track) F.sn[row_index(en, i)] = sn_init(sc, c0.S, c0.sl, en.contig_idx, n);
                cm_add(part, c0.S, c0.sl, i);
            }
            if (!(part.S == 0 && part.row == 0)) throw Error(STITCH_ERR_INTERNAL, "emul: column-0 best is not (0, row 0)");
            cm[a] = 0; cmk[a] = 0; cml[a] = 0;
            Col0 cmm = col0_at(sc, en.m, en.m);
            Sm[a] = cmm.S; slm[a] = cmm.sl; tbm[a] = cmm.s_tb;
            int32_t t; uint32_t lx; col0_tracker(sc, en.m, t, lx);
            F.colrec[a].lx = lx;
        }
        std::vector<JumpInfo> J(C);
        for (uint32_t j = 1; j <= n; ++j) {
            int32_t g = cm[0];
            for (uint32_t a = 1; a < C; ++a) g = std::max(g, cm[a]);
            F.gcol[j - 1] = g;
            for (uint32_t a = 0; a < C; ++a) J[a] = select_jump(sc, L.ent.data(), C, a, cm.data(), cml.data(), cmk.data());
            const bool ck = (j % K == 0) && j < n;
            ColumnArgs A{};
            A.sc = &sc; A.ent = L.ent.data(); A.C = C; A.n_tiles = L.n_tiles; A.bases = bases; A.read = job.read;
            A.j = j; A.n = n; A.prev = st[(j - 1) & 1].data(); A.curr = st[j & 1].data();
            A.ck_state = ck ? F.ck_state.data() + (size_t)(j / K - 1) * PM : nullptr;
            A.tb_col = nullptr; A.colrec_col = F.colrec.data() + (size_t)j * C; A.J = J.data();
            A.Sm = Sm.data(); A.slm = slm.data(); A.tbm = tbm.data(); A.cm = cm.data(); A.cml = cml.data(); A.cmk = cmk.data();
            A.track = j >= track_from; A.sn = F.sn.data(); A.lastcol = j == n; A.last = F.last.data();
            emul_column(A);
            if (ck) for (uint32_t a = 0; a < C; ++a) F.ck_sum[(size_t)(j / K - 1) * C + a] = CkSum{Sm[a], slm[a], tbm[a], 0};
        }
        {
            int32_t g = cm[0];
            for (uint32_t a = 1; a < C; ++a) g = std::max(g, cm[a]);
            F.gcol[n] = g;
        }
        F.need_full_track = any_track && track_from > 1 && first_candidate_column(sc, F.gcol.data(), n) < track_from;
    }

    // Re-fills contig `a` over the block of columns holding column j.
    void load_unit(const Job &job, const Layout &L, const Fill &F, uint32_t a, uint32_t j, std::vector<uint8_t> &bytes, TbUnit &u) {
        const Scoring &sc = al.opts.sc;
        const uint32_t C = (uint32_t)L.ent.size(), PM = L.PM(), n = job.n;
        const uint32_t b = (j - 1) / K, jb = b * K, je = std::min(jb + K, n);
        ContigEntry en = L.ent[a];
        const uint32_t pm = en.ntiles * TILE, gbase = en.tile_start * TILE;
        en.tile_start = 0;
        std::vector<CellState> st[2];
        st[0].assign(pm, CellState{MIN_SCORE, MIN_SCORE, 0, 0}); st[1] = st[0];
        int32_t Sm; uint32_t slm, tbm;
        if (b == 0) {
            for (uint32_t i = 1; i <= en.m; ++i) { Col0 c0 = col0_at(sc, i, en.m); st[0][row_index(en, i)] = CellState{c0.S, MIN_SCORE, c0.sl, 0}; }
            Col0 cmm = col0_at(sc, en.m, en.m);
            Sm = cmm.S; slm = cmm.sl; tbm = cmm.s_tb;
        } else {
            const CellState *ck = F.ck_state.data() + (size_t)(b - 1) * PM + gbase;
            for (uint32_t p = 0; p < pm; ++p) st[jb & 1][p] = ck[p];
            const CkSum &cs = F.ck_sum[(size_t)(b - 1) * C + a];
            Sm = cs.Sm; slm = cs.slm; tbm = cs.tbm;
        }
        bytes.assign((size_t)(je - jb) * pm, 0);
        stats.cells += (uint64_t)en.m * (je - jb);
        for (uint32_t jj = jb + 1; jj <= je; ++jj) {
            const ColRec &cr = F.colrec[(size_t)jj * C + a];
            JumpInfo J{cr.jscore, cr.jlen, cr.jidx, cr.jfrom};
            ColumnArgs A{};
            A.sc = &sc; A.ent = &en; A.C = 1; A.n_tiles = en.ntiles; A.bases = al.contigs.blob.data(); A.read = job.read;
            A.j = jj; A.n = n; A.prev = st[(jj - 1) & 1].data(); A.curr = st[jj & 1].data(); A.ck_state = nullptr;
            A.tb_col = bytes.data() + (size_t)(jj - jb - 1) * pm; A.colrec_col = nullptr; A.J = &J;
            A.Sm = &Sm; A.slm = &slm; A.tbm = &tbm; A.cm = nullptr; A.cml = nullptr; A.cmk = nullptr;
            A.track = false; A.sn = nullptr; A.lastcol = false; A.last = nullptr;
            emul_column(A);
        }
        u.bytes = bytes.data(); u.a = a; u.jb = jb; u.je = je; u.pm = pm;
    }

    void run_one(const Job &job, JobResult &res) {
        const Layout &L = al.layouts.layouts[job.layout];
        const Scoring &sc = al.opts.sc;
        const uint32_t C = (uint32_t)L.ent.size(), n = job.n;
        if (n == 0) throw Error(STITCH_ERR_INVALID, "empty read");
        const bool tracked_mode = sc.ys != MIN_SCORE;   // Sn can only matter when y-suffix clipping is free
        uint32_t track_from = tracked_mode ? (n > WINDOW ? n - WINDOW + 1 : 1) : n + 1;
        Fill F;
        fill(job, L, track_from, F);
        if (F.need_full_track) { ++stats.launches; fill(job, L, 1, F); }
        for (uint32_t a = 0; a < C; ++a)
            fixup_contig(sc, L.ent[a], n, F.last.data(), F.sn.data(), tracked_mode, &F.colrec[(size_t)n * C + a].lx);
        if (const char *dump = std::getenv("STITCH_DUMP_DIR")) dump_job(dump, dump_seq++, F.last, F.sn, F.colrec, std::vector<uint8_t>());
        ReadView v;
        v.sc = sc; v.ent = L.ent.data(); v.C = C; v.n = n; v.colrec = F.colrec.data();
        v.last = F.last.data(); v.sn = F.sn.data(); v.contig_bases = al.contigs.blob.data(); v.read = job.read;
        v.pos_of = L.pos_of.data();
        v.unit.bytes = nullptr; v.unit.a = 0xffffffffu; v.unit.jb = v.unit.je = v.unit.pm = 0;
        std::vector<uint8_t> unit_bytes;
        const uint32_t cap = 2 * n + 4 * C + 64;
        auto do_walk = [&](uint32_t a_end, RawChain &rc) -> uint32_t {
            uint32_t c = cap;
            for (;;) {
                rc.ops.assign(c, OutOp{0, 0, 0});
                WalkState ws;
                walk_begin(v, a_end, rc.ops.data(), c, ws, rc.h);
                uint32_t s;
                while ((s = walk_run(v, ws, rc.h)) == WALK_NEED_UNIT) load_unit(job, L, F, ws.a, ws.j, unit_bytes, v.unit);
                if (s != WALK_OVERFLOW) break;
                c *= 2;
            }
            rc.ops.resize(rc.h.n_ops);
            return rc.h.status;
        };
        if (job.walk == WALK_BEST) {
            RawChain rc;
            uint32_t s = do_walk(pick_end(v, nullptr), rc);
            if (s == WALK_PANIC) throw Error(STITCH_ERR_INTERNAL, "traceback reached a state the reference panics on");
            if (s == WALK_OK) res.chains.push_back(std::move(rc));
        } else if (job.walk == WALK_FROM) {
            const int16_t a_end = L.pos_of[job.from_contig];
            if (a_end >= 0) {
                RawChain rc;
                uint32_t s = do_walk((uint32_t)a_end, rc);
                if (s == WALK_PANIC) throw Error(STITCH_ERR_INTERNAL, "traceback reached a state the reference panics on");
                if (s == WALK_OK) res.chains.push_back(std::move(rc));
            }
        } else {   // traceback_all (TB:152-217) over the layout's contigs
            std::vector<uint8_t> seen(C, 0);
            uint32_t n_seen = 0;
            auto mark = [&](uint32_t idx) {
                const int16_t p = idx < MAX_STRANDS ? L.pos_of[idx] : (int16_t)-1;
                if (p >= 0 && !seen[(size_t)p]) { seen[(size_t)p] = 1; ++n_seen; }
            };
            while (n_seen < C) {
                const uint32_t a_end = pick_end(v, seen.data());
                RawChain rc;
                uint32_t s = do_walk(a_end, rc);
                if (s == WALK_PANIC) throw Error(STITCH_ERR_INTERNAL, "traceback reached a state the reference panics on");
                if (s != WALK_OK) { mark(L.ent[a_end].contig_idx); continue; }
                mark(rc.h.start_contig_idx); mark(rc.h.end_contig_idx);
                for (const OutOp &o : rc.ops) if (o.kind == OP_XJUMP) mark(o.a);
                res.chains.push_back(std::move(rc));
            }
        }
    }

    void run(const std::vector<Job> &jobs, std::vector<JobResult> &out) override {
        out.assign(jobs.size(), JobResult());
        for (size_t k = 0; k < jobs.size(); ++k) run_one(jobs[k], out[k]);
    }
};

Backend *stitch_make_backend(Aligner &al, int) { return new EmulBackend(al); }

}  // namespace host
}  // namespace stitch
