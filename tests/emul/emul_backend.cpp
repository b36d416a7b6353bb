// TEST INFRASTRUCTURE ONLY — CPU emulation of the CUDA fill kernel's tile / lane / round
// structure over the product's own __host__ __device__ DP core (stitch_b200/csrc/dp_core.h) and
// host driver (host_common.hpp).  It lets the CPU-only test tier fuzz the decomposition (pass A,
// insertion-chain scan, pass B, row-m finalize, fix-up, walk, re-alignment driver) against the
// oracle.  Exported under the emul_ prefix; the product library never links or loads this.
#define STITCH_API(name) emul_##name
#include "../../stitch_b200/csrc/capi_impl.hpp"

#include <vector>

namespace stitch {
namespace host {

#ifndef EMUL_WARPS
#define EMUL_WARPS 4
#endif

struct EmulBackend : Backend {
    Aligner &al;
    uint32_t dump_seq = 0;
    explicit EmulBackend(Aligner &a) : al(a) {}

    void run_one(const Job &job, JobResult &res) {
        const Layout &L = al.layouts.layouts[job.layout];
        const Scoring &sc = al.opts.sc;
        const uint32_t C = (uint32_t)L.ent.size(), PM = L.PM(), n = job.n;
        const uint8_t *bases = al.contigs.blob.data();
        const bool track = sc.ys != MIN_SCORE;   // Sn can only matter when y-suffix clipping is free
        std::vector<CellState> st[2];
        st[0].assign(PM, CellState{MIN_SCORE, MIN_SCORE, 0, 0});
        st[1] = st[0];
        std::vector<uint8_t> tb((size_t)std::max<uint32_t>(n, 1) * PM, 0);
        std::vector<ColRec> colrec((size_t)(n + 1) * C, ColRec{0, 0, 0, 0});
        std::vector<LastCell> last(PM);
        std::vector<SnRec> sn(PM, SnRec{track ? MIN_SCORE : 0x3fffffff, 7, 7, 7});
        stats.cells += (uint64_t)L.cells_per_col * n;
        stats.fills += 1;

        // per-contig summaries of the previous column
        std::vector<int32_t> cm(C), Sm(C);
        std::vector<uint32_t> cml(C), cmk(C), slm(C), tbm(C);
        // column 0
        for (uint32_t a = 0; a < C; ++a) {
            const ContigEntry &en = L.ent[a];
            CmPart part; cm_init(part); cm_add(part, 0, 0, 0);
            for (uint32_t i = 1; i <= en.m; ++i) {
                Col0 c0 = col0_at(sc, i, en.m);
                const uint32_t r = i - 1, tile = en.tile_start + r / TILE, lane = (r % TILE) / STRIP, k = r % STRIP;
                st[0][state_index(tile, lane, k)] = CellState{c0.S, MIN_SCORE, c0.sl, 0};
                sn[state_index(tile, lane, k)] = sn_init(sc, c0.S, c0.sl, en.contig_idx, n);
                cm_add(part, c0.S, c0.sl, i);
                if (n == 0) {   // column 0 is also column n
                    LastCell lc{}; lc.S = c0.S; lc.I = c0.I; lc.sl = c0.sl; lc.il = c0.il; lc.idx = en.contig_idx; lc.from = 0;
                    lc.s_tb = (uint8_t)c0.s_tb; lc.i_tb = (uint8_t)c0.i_tb; lc.flags = 0x80;
                    last[state_index(tile, lane, k)] = lc;
                }
            }
            if (!(part.S == 0 && part.row == 0)) throw Error(STITCH_ERR_INTERNAL, "emul: column-0 best is not (0, row 0)");
            cm[a] = 0; cmk[a] = 0; cml[a] = 0;
            Col0 cmm = col0_at(sc, en.m, en.m);
            Sm[a] = cmm.S; slm[a] = cmm.sl; tbm[a] = cmm.s_tb;
            int32_t t; uint32_t lx; col0_tracker(sc, en.m, t, lx);
            colrec[a].lx = lx;
        }
        if (n == 0) throw Error(STITCH_ERR_INVALID, "empty read");

        const int W = EMUL_WARPS;
        std::vector<LaneA> la((size_t)W * 32);
        std::vector<ICarry> excl((size_t)W * 32), tileagg((size_t)W);
        std::vector<ICarry> tcarry((size_t)W);
        std::vector<JumpInfo> J(C);
        std::vector<XsPart> xs(C); std::vector<CmPart> cmp(C); std::vector<RowM> rowm(C);

        for (uint32_t j = 1; j <= n; ++j) {
            const std::vector<CellState> &prev = st[(j - 1) & 1];
            std::vector<CellState> &curr = st[j & 1];
            const Row0 r0 = row0_at(sc, j, n), r0p = row0_at(sc, j - 1, n);
            ColConst cc; cc.j = j; cc.n = n; cc.q = job.read[j - 1];
            cc.xclip_score = sc.xp + std::max(sc.yp, sc.o + sc.e * (int32_t)j);
            cc.sl0j = r0.sl;
            uint8_t *tb_col = tb.data() + (size_t)(j - 1) * PM;
            for (uint32_t a = 0; a < C; ++a) {
                J[a] = select_jump(sc, L.ent.data(), C, a, cm.data(), cml.data(), cmk.data());
                xs_init(xs[a]); cm_init(cmp[a]);
            }
            ICarry round_carry{MIN_SCORE, 0, 0};
            for (uint32_t t0 = 0; t0 < L.n_tiles; t0 += (uint32_t)W) {
                const int nw = (int)std::min<uint32_t>((uint32_t)W, L.n_tiles - t0);
                std::vector<TileCtx> tcs((size_t)nw);
                // phase A
                for (int w = 0; w < nw; ++w) {
                    const uint32_t tile = t0 + (uint32_t)w;
                    uint32_t a = 0;
                    while (!(tile >= L.ent[a].tile_start && tile < L.ent[a].tile_start + L.ent[a].ntiles)) ++a;
                    const ContigEntry &en = L.ent[a];
                    TileCtx tc; tc.a = a; tc.self_idx = en.contig_idx; tc.m = en.m; tc.tile = tile;
                    tc.tile_in_contig = tile - en.tile_start; tc.J = J[a]; tc.circular = en.circular != 0;
                    tc.wrap_src_ok = tbm[a] != TB_XCLIP_SUFFIX; tc.Sm_prev = Sm[a]; tc.slm_prev = slm[a];
                    tcs[(size_t)w] = tc;
                    for (uint32_t lane = 0; lane < 32; ++lane) {
                        const uint32_t row0 = tc.tile_in_contig * TILE + lane * STRIP + 1;
                        LaneA &out = la[(size_t)w * 32 + lane];
                        out.has_m = 0; out.agg = ICarry{MIN_SCORE, 0, 0};
                        if (row0 > en.m) continue;
                        CellState up[STRIP]; uint8_t x[STRIP];
                        for (int k = 0; k < STRIP; ++k) {
                            up[k] = prev[state_index(tile, lane, (uint32_t)k)];
                            const uint32_t i = row0 + (uint32_t)k;
                            x[k] = i <= en.m ? bases[en.seq_off + i - 1] : 0;
                        }
                        int32_t dgS; uint32_t dgsl;
                        if (row0 == 1) { dgS = r0p.S; dgsl = r0p.sl; }
                        else {
                            const uint32_t r = row0 - 2, pt = en.tile_start + r / TILE, pl = (r % TILE) / STRIP, pk = r % STRIP;
                            dgS = prev[state_index(pt, pl, pk)].S; dgsl = prev[state_index(pt, pl, pk)].sl;
                        }
                        lane_pass_a(sc, cc, tc, row0, up, dgS, dgsl, x, out, &rowm[a]);
                    }
                    // warp scan (Hillis-Steele over lane aggregates, as the kernel does with shuffles)
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
                // carry folding within the round (each warp folds the aggregates of the tiles before it)
                for (int w = 0; w < nw; ++w) {
                    const TileCtx &tc = tcs[(size_t)w];
                    // walk back to the first tile of this contig inside the round
                    int w0 = w;
                    while (w0 > 0 && tcs[(size_t)(w0 - 1)].a == tc.a) --w0;
                    ICarry c;
                    if (tcs[(size_t)w0].tile_in_contig == 0) c = icarry_row1(sc, r0);
                    else c = round_carry;   // carry into the round's first tile (same contig continues)
                    for (int u = w0; u < w; ++u) c = icarry_combine(c, TILE, sc.e, tileagg[(size_t)u]);
                    tcarry[(size_t)w] = c;
                }
                // carry out of the round
                {
                    const int w = nw - 1;
                    round_carry = icarry_combine(tcarry[(size_t)w], TILE, sc.e, tileagg[(size_t)w]);
                }
                // phase B
                for (int w = 0; w < nw; ++w) {
                    const TileCtx &tc = tcs[(size_t)w];
                    const ContigEntry &en = L.ent[tc.a];
                    for (uint32_t lane = 0; lane < 32; ++lane) {
                        const uint32_t row0 = tc.tile_in_contig * TILE + lane * STRIP + 1;
                        if (row0 > en.m) continue;
                        LaneA &a_ = la[(size_t)w * 32 + lane];
                        ICarry cin = lane == 0 ? tcarry[(size_t)w]
                                               : icarry_combine(tcarry[(size_t)w], lane * STRIP, sc.e, excl[(size_t)w * 32 + lane]);
                        uint8_t x[STRIP];
                        for (int k = 0; k < STRIP; ++k) {
                            const uint32_t i = row0 + (uint32_t)k;
                            x[k] = i <= en.m ? bases[en.seq_off + i - 1] : 0;
                        }
                        LaneB lb;
                        lane_pass_b(sc, cc, tc, row0, lane, a_, cin, curr.data(), tb_col, track, sn.data(), j == n,
                                    last.data(), x, lb, &rowm[tc.a]);
                        xs[tc.a] = xs_merge(xs[tc.a], lb.xs);
                        cmp[tc.a] = cm_merge(cmp[tc.a], lb.cm);
                    }
                }
            }
            // finalize every contig
            for (uint32_t a = 0; a < C; ++a) {
                ContigColOut o = contig_finalize(sc, cc, L.ent[a], a, C, rowm[a], xs[a], cmp[a], r0, J[a], curr.data(), tb_col,
                                                 colrec.data() + (size_t)j * C, track, sn.data(), j == n, last.data());
                cm[a] = o.cm.S; cmk[a] = o.cm.row; cml[a] = o.cm.sl;
                Sm[a] = o.Sm; slm[a] = o.slm; tbm[a] = o.s_tb_m;
            }
        }
        // end-of-read fix-up, then the walks
        for (uint32_t a = 0; a < C; ++a)
            fixup_contig(sc, L.ent[a], n, last.data(), sn.data(), track, &colrec[(size_t)n * C + a].lx);
        if (const char *dump = std::getenv("STITCH_DUMP_DIR")) dump_job(dump, dump_seq++, last, sn, colrec, tb);
        ReadView v;
        v.sc = sc; v.ent = L.ent.data(); v.C = C; v.n = n; v.PM = PM; v.tb = tb.data(); v.colrec = colrec.data();
        v.last = last.data(); v.sn = sn.data(); v.contig_bases = bases; v.read = job.read; v.pos_of = L.pos_of.data();
        const uint32_t cap = 2 * n + 4 * C + 64;   // RLE ops never exceed this for sane chains; grown on overflow
        auto do_walk = [&](uint32_t a_end, RawChain &rc) -> uint32_t {
            uint32_t c = cap;
            for (;;) {
                rc.ops.assign(c, OutOp{0, 0, 0});
                walk_chain(v, a_end, rc.ops.data(), c, rc.h);
                if (rc.h.status != WALK_OVERFLOW) break;
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
