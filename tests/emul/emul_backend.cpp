// TEST INFRASTRUCTURE ONLY — CPU emulation of the CUDA kernels' tile / lane / round structure
// over the product's own __host__ __device__ DP core (stitch_b200/csrc/dp_core.h) and host driver
// (host_common.hpp).  It lets the CPU-only test tier fuzz the decomposition (pass A, insertion-chain
// scan, pass B, row-m finalize, checkpoints, windowed y-suffix tracking, fix-up, unit re-fill +
// resumable walk, re-alignment driver) against the oracle.  Exported under the emul_ prefix; the
// product library never links or loads this.
#define STITCH_API(name) emul_##name
#include "../../stitch_b200/csrc/capi_impl.hpp"
#include "../../stitch_b200/csrc/dp_packed.h"

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
    uint32_t K, WINDOW, PACKED;
    explicit EmulBackend(Aligner &a) : al(a) {
        K = std::max<uint32_t>(1, env_u32("EMUL_K", 7));          // checkpoint spacing (columns)
        WINDOW = std::max<uint32_t>(1, env_u32("EMUL_WINDOW", 6));  // columns at the end of the read filled by the wide path
        PACKED = env_u32("EMUL_PACKED", 1);                         // 0: wide path only
    }

    struct Fill {
        std::vector<ColRec> colrec; std::vector<LastCell> last; std::vector<SnRec> sn;
        std::vector<CellState> ck_state; std::vector<CkSum> ck_sum; std::vector<int32_t> gcol;
        std::vector<CellState> hand_state; std::vector<CkSum> hand_sum;   // wide state at column j0 (packed -> wide hand-over)
        bool need_full_track = false;
    };

    void alloc(const Job &job, const Layout &L, Fill &F) {
        const uint32_t C = (uint32_t)L.ent.size(), PM = L.PM(), n = job.n;
        const uint32_t nb = (n + K - 1) / K;
        F.colrec.assign((size_t)(n + 1) * C, ColRec{});
        F.last.assign(PM, LastCell{});
        F.sn.assign(PM, SnRec{0x3fffffff, 7, 7, 7});   // garbage unless initialised below
        F.ck_state.assign((size_t)(nb ? nb - 1 : 0) * PM, CellState{});
        F.ck_sum.assign((size_t)(nb ? nb - 1 : 0) * C, CkSum{});
        F.gcol.assign((size_t)n + 1, MIN_SCORE);
        F.hand_state.assign(PM, CellState{MIN_SCORE, MIN_SCORE, 0, 0});
        F.hand_sum.assign(C, CkSum{});
        for (uint32_t a = 0; a < C; ++a) { int32_t t; uint32_t lx; col0_tracker(al.opts.sc, L.ent[a].m, t, lx); F.colrec[a].lx = lx; }
    }

    // Wide columns (j0, n].  j0 == 0: from column 0; else from the hand-over state at column j0 (the jump of
    // column j0 + 1 is read from colrec).  gcol[0..j0] must be filled by the caller when j0 > 0.
    void fill_wide(const Job &job, const Layout &L, uint32_t j0, uint32_t track_from, Fill &F) {
        const Scoring &sc = al.opts.sc;
        const uint32_t C = (uint32_t)L.ent.size(), PM = L.PM(), n = job.n;
        const uint8_t *bases = al.contigs.blob.data();
        std::vector<CellState> st[2];
        st[0].assign(PM, CellState{MIN_SCORE, MIN_SCORE, 0, 0}); st[1] = st[0];
        stats.cells += (uint64_t)L.cells_per_col * (n - j0);
        stats.fills += 1;
        const bool any_track = track_from <= n;
        std::vector<int32_t> cm(C), Sm(C);
        std::vector<uint32_t> cml(C), cmk(C), slm(C), tbm(C);
        for (uint32_t a = 0; a < C; ++a) {
            const ContigEntry &en = L.ent[a];
            CmPart part; cm_init(part); cm_add(part, 0, 0, 0);
            for (uint32_t i = 1; i <= en.m; ++i) {
                Col0 c0 = col0_at(sc, i, en.m);
                if (j0 == 0) st[0][row_index(en, i)] = CellState{c0.S, MIN_SCORE, c0.sl, 0};
                if (any_track) F.sn[row_index(en, i)] = sn_init(sc, c0.S, c0.sl, en.contig_idx, n);
                cm_add(part, c0.S, c0.sl, i);
            }
            if (!(part.S == 0 && part.row == 0)) throw Error(STITCH_ERR_INTERNAL, "emul: column-0 best is not (0, row 0)");
            cm[a] = 0; cmk[a] = 0; cml[a] = 0;
            Col0 cmm = col0_at(sc, en.m, en.m);
            Sm[a] = cmm.S; slm[a] = cmm.sl; tbm[a] = cmm.s_tb;
        }
        if (j0 > 0) {
            st[j0 & 1] = F.hand_state;
            for (uint32_t a = 0; a < C; ++a) { Sm[a] = F.hand_sum[a].Sm; slm[a] = F.hand_sum[a].slm; tbm[a] = F.hand_sum[a].tbm; }
        }
        std::vector<JumpInfo> J(C);
        for (uint32_t j = j0 + 1; j <= n; ++j) {
            if (j == j0 + 1 && j0 > 0) {
                for (uint32_t a = 0; a < C; ++a) { const ColRec &cr = F.colrec[(size_t)j * C + a]; J[a] = JumpInfo{cr.jscore, cr.jlen, cr.jidx, cr.jfrom}; }
            } else {
                int32_t g = cm[0];
                for (uint32_t a = 1; a < C; ++a) g = std::max(g, cm[a]);
                F.gcol[j - 1] = g;
                for (uint32_t a = 0; a < C; ++a) J[a] = select_jump(sc, L.ent.data(), C, a, cm.data(), cml.data(), cmk.data());
            }
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

    // Packed columns [1, j1] (dp_packed.h) with the kernel's structure: EMUL_WARPS warps own contiguous
    // chunks of tiles; a chunk that starts inside a contig recomputes the chain exit of the strip before it
    // (the halo).  Leaves wide checkpoints, the wide hand-over state at column j1, colrec[1..j1+1], gcol[0..j1].
    void fill_packed(const Job &job, const Layout &L, uint32_t LB, uint32_t j1, Fill &F) {
        const Scoring &sc = al.opts.sc;
        const uint32_t C = (uint32_t)L.ent.size(), PM = L.PM(), NT = L.n_tiles, n = job.n;
        const uint8_t *bases = al.contigs.blob.data();
        const PK pk = pk_make(sc, LB);
        const int W = EMUL_WARPS;
        stats.cells += (uint64_t)L.cells_per_col * j1;
        std::vector<int32_t> Sst[2], Dst[2];   // linear: tile * TILE + (row - 1 within the contig's tiles)
        for (int b = 0; b < 2; ++b) { Sst[b].assign(PM, pk.NEGKEY); Dst[b].assign(PM, pk.NEGKEY + pk.PD6); }
        std::vector<uint32_t> owner(NT);
        for (uint32_t a = 0; a < C; ++a) for (uint32_t t = 0; t < L.ent[a].ntiles; ++t) owner[L.ent[a].tile_start + t] = a;
        std::vector<int32_t> cm(C, 0), Sm(C);
        std::vector<uint32_t> cml(C, 0), cmk(C, 0), slm(C), tbm(C);
        std::vector<int32_t> SmKey(C);
        int32_t B = 0;   // B_0 = 0
        for (uint32_t a = 0; a < C; ++a) {
            const ContigEntry &en = L.ent[a];
            for (uint32_t i = 1; i <= en.m; ++i) {
                Col0 c0 = col0_at(sc, i, en.m);
                Sst[0][row_linear(en, i)] = pk_from_wide(pk, 0, c0.S, c0.sl, 0);
            }
            Col0 cmm = col0_at(sc, en.m, en.m);
            Sm[a] = cmm.S; slm[a] = cmm.sl; tbm[a] = cmm.s_tb;
            SmKey[a] = Sst[0][row_linear(en, en.m)];
        }
        F.gcol[0] = 0;
        std::vector<JumpInfo> J(C);
        std::vector<int32_t> tilemax(NT);
        struct RowMStash { int32_t diag, D6, jp, I; bool wrap; };
        std::vector<RowMStash> stash(C);
        for (uint32_t j = 1; j <= j1 + 1; ++j) {
            int32_t g = cm[0];
            for (uint32_t a = 1; a < C; ++a) g = std::max(g, cm[a]);
            const int32_t Bprev = B;
            B = g;
            for (uint32_t a = 0; a < C; ++a) {
                J[a] = select_jump(sc, L.ent.data(), C, a, cm.data(), cml.data(), cmk.data());
                ColRec &cr = F.colrec[(size_t)j * C + a];
                cr.jscore = J[a].score; cr.jlen = J[a].len; cr.jidx = J[a].idx; cr.jfrom = J[a].from;
            }
            if (j == j1 + 1) break;   // only the jump of the first wide column was needed
            const PCol pc = pk_col(pk, sc, B, Bprev, j, n, job.read[j - 1]);
            const Row0 r0 = row0_at(sc, j, n), r0p = row0_at(sc, j - 1, n);
            const int32_t *Sp = Sst[(j - 1) & 1].data(), *Dp = Dst[(j - 1) & 1].data();
            int32_t *Sc = Sst[j & 1].data(), *Dc = Dst[j & 1].data();
            for (int w = 0; w < W; ++w) {
                const uint32_t t_lo = (uint32_t)((uint64_t)NT * w / W), t_hi = (uint32_t)((uint64_t)NT * (w + 1) / W);
                int32_t prev_exit = 0;
                for (uint32_t tile = t_lo; tile < t_hi; ++tile) {
                    const uint32_t a = owner[tile];
                    const ContigEntry &en = L.ent[a];
                    const uint32_t tic = tile - en.tile_start;
                    const bool first = tic == 0, lastt = tic + 1 == en.ntiles;
                    const bool special = first || lastt;
                    const int32_t Jc = pk_jc(pk, pc, J[a].score, J[a].len);
                    const bool wrap0 = first && en.circular && tbm[a] != TB_XCLIP_SUFFIX;
                    const int32_t wbase = pk_wbase(pk, SmKey[a]);
                    PStrip strips[32];
                    int nvs[32]; bool hasm[32];
                    for (uint32_t lane = 0; lane < 32; ++lane) {
                        const uint32_t row0 = tic * TILE + lane * STRIP + 1;   // 1-based row of the strip's first cell
                        const uint32_t base = tile * TILE + lane * STRIP;
                        int32_t Sdg0;
                        if (row0 == 1) Sdg0 = pk_from_wide(pk, Bprev, r0p.S, r0p.sl, 0);
                        else Sdg0 = Sp[base - 1];
                        uint8_t x[STRIP];
                        for (int k = 0; k < STRIP; ++k) x[k] = row0 + k <= en.m ? bases[en.seq_off + row0 + k - 1] : 0;
                        int nv = STRIP; bool hm = false;
                        if (special) {
                            const int64_t left = (int64_t)en.m - (int64_t)row0;   // rows < m in this strip
                            nv = left >= STRIP ? STRIP : (left < 0 ? 0 : (int)left);
                            hm = left >= 0 && left < STRIP;
                            for (int k = 0; k < STRIP; ++k) {
                                strips[lane].YC[k] = pk.NEGKEY;
                                const uint32_t i = row0 + (uint32_t)k;
                                if (first && sc.yp != MIN_SCORE && sc.xp == MIN_SCORE && i <= en.m)
                                    strips[lane].YC[k] = pk_key(pk, (int64_t)sc.yp + sc.o + (int64_t)sc.e * i - B, PP_YC, col0_slen(sc, i, en.m));
                            }
                        }
                        nvs[lane] = nv; hasm[lane] = hm;
                        if (special) pk_pass1<true, false>(pk, pc, Sp + base, Dp + base, Sdg0, x, Jc, wrap0 && lane == 0, wbase, nv, hm, strips[lane]);
                        else pk_pass1<false, false>(pk, pc, Sp + base, Dp + base, Sdg0, x, Jc, false, wbase, STRIP, false, strips[lane]);
                    }
                    int32_t tmax = pk.NEGKEY;
                    for (uint32_t lane = 0; lane < 32; ++lane) {
                        const uint32_t base = tile * TILE + lane * STRIP;
                        int32_t cin;
                        if (lane > 0) cin = pk_carry_from_exit(pk, strips[lane - 1].exit);
                        else if (first) cin = pk_carry_row1(pk, pc, sc, r0);
                        else if (tile == t_lo) {   // chunk start inside a contig: the halo strip (rows row0-8 .. row0-1)
                            const uint32_t hb = tile * TILE - STRIP;
                            uint8_t x[STRIP];
                            const uint32_t hrow0 = tic * TILE - STRIP + 1;
                            for (int k = 0; k < STRIP; ++k) x[k] = bases[en.seq_off + hrow0 + k - 1];
                            const bool hfirst = (tic == 1);   // the halo lies in the contig's first tile
                            PStrip h;
                            for (int k = 0; k < STRIP; ++k) {
                                h.YC[k] = pk.NEGKEY;
                                const uint32_t i = hrow0 + (uint32_t)k;
                                if (hfirst && sc.yp != MIN_SCORE && sc.xp == MIN_SCORE)
                                    h.YC[k] = pk_key(pk, (int64_t)sc.yp + sc.o + (int64_t)sc.e * i - B, PP_YC, col0_slen(sc, i, en.m));
                            }
                            pk_pass1<true, false>(pk, pc, Sp + hb, Dp + hb, Sp[hb - 1], x, Jc, false, wbase, STRIP, false, h);
                            cin = pk_carry_from_exit(pk, h.exit);
                        } else cin = pk_carry_from_exit(pk, prev_exit);
                        int32_t S[STRIP]; int32_t colmax = pk.NEGKEY; int32_t I_m = pk.NEGKEY; uint32_t iext_m = 0;
                        if (special) pk_pass2<true, false>(pk, pc, strips[lane], cin, 0, nvs[lane], hasm[lane], S, colmax, nullptr, I_m, iext_m);
                        else pk_pass2<false, false>(pk, pc, strips[lane], cin, 0, STRIP, false, S, colmax, nullptr, I_m, iext_m);
                        for (int k = 0; k < nvs[lane]; ++k) { Sc[base + k] = S[k]; Dc[base + k] = strips[lane].D6[k]; }
                        if (hasm[lane]) {
                            const int km = nvs[lane];
                            stash[a] = RowMStash{strips[lane].A[km], strips[lane].D6[km], strips[lane].jp[km], I_m, false};
                            Dc[base + km] = strips[lane].D6[km];
                        }
                        tmax = pk_max(tmax, colmax);
                    }
                    tilemax[tile] = tmax;
                    prev_exit = strips[31].exit;
                }
            }
            // per contig: tracker / column best over rows < m, finish row m (wide arithmetic), column best
            for (uint32_t a = 0; a < C; ++a) {
                const ContigEntry &en = L.ent[a];
                int32_t kmax = pk.NEGKEY;
                for (uint32_t t = 0; t < en.ntiles; ++t) kmax = pk_max(kmax, tilemax[en.tile_start + t]);
                CmPart rows; cm_init(rows);
                XsPart tr; xs_init(tr);
                if (en.m >= 2) {
                    const int32_t smax = pk_rel(pk, kmax);
                    uint32_t frow = 0;
                    for (uint32_t t = 0; t < en.ntiles && !frow; ++t) {
                        if (pk_rel(pk, tilemax[en.tile_start + t]) != smax) continue;
                        for (uint32_t r = 0; r < (uint32_t)TILE; ++r) {
                            const uint32_t i = t * TILE + r + 1;
                            if (i >= en.m) break;
                            if (pk_rel(pk, Sc[(en.tile_start + t) * TILE + r]) == smax) { frow = i; break; }
                        }
                    }
                    if (!frow) throw Error(STITCH_ERR_INTERNAL, "emul packed: column best not found");
                    const int32_t fkey = Sc[row_linear(en, frow)];
                    rows.S = B + smax; rows.row = frow; rows.sl = pk_len(pk, fkey); rows.valid = 1;
                    if (sc.xs != MIN_SCORE) { tr.t = B + smax + sc.xs; tr.len = pk_len(pk, kmax); tr.row = 1; }
                }
                const RowMStash &sm_ = stash[a];
                RowM rm{};
                rm.diag = pk_abs(pk, B, sm_.diag); rm.dgl = pk_len(pk, sm_.diag);
                rm.D = pk_abs(pk, B, sm_.D6); rm.dl = pk_len(pk, sm_.D6); rm.dext = 0;
                rm.I = pk_abs(pk, B, sm_.I); rm.il = pk_len(pk, sm_.I); rm.iext = 0;
                rm.jp.score = pk_abs(pk, B, sm_.jp); rm.jp.len = pk_len(pk, sm_.jp); rm.jp.idx = 0; rm.jp.from = 0;
                rm.xclip = sc.xp + std::max(sc.yp, sc.o + sc.e * (int32_t)j); rm.xclip_len = r0.sl;
                rm.yclip = sc.yp + sc.o + sc.e * (int32_t)en.m; rm.yclip_len = 0;
                rm.is_match = bases[en.seq_off + en.m - 1] == pc.q;
                const RowMOut ro = finish_rowm(sc, rm, tr, en.contig_idx, en.m);
                Sc[row_linear(en, en.m)] = pk_from_wide(pk, B, ro.c.S, ro.c.sl, 0);
                CmPart cmv; cm_init(cmv);
                cm_add(cmv, r0.S, r0.sl, 0);
                cmv = cm_merge(cmv, rows);
                CmPart top; top.S = ro.c.S; top.row = en.m; top.sl = ro.c.sl; top.valid = 1;
                cmv = cm_merge(cmv, top);
                cm[a] = cmv.S; cmk[a] = cmv.row; cml[a] = cmv.sl;
                Sm[a] = ro.c.S; slm[a] = ro.c.sl; tbm[a] = ro.s_tb;
                SmKey[a] = Sc[row_linear(en, en.m)];
            }
            {
                int32_t gg = cm[0];
                for (uint32_t a = 1; a < C; ++a) gg = std::max(gg, cm[a]);
                F.gcol[j] = gg;
            }
            // wide checkpoints / hand-over
            const bool ck = (j % K == 0) && j < n;
            if (ck || j == j1) {
                for (uint32_t a = 0; a < C; ++a) {
                    const ContigEntry &en = L.ent[a];
                    for (uint32_t i = 1; i <= en.m; ++i) {
                        const int32_t s = Sc[row_linear(en, i)], d = Dc[row_linear(en, i)];
                        const CellState cs{pk_abs(pk, B, s), pk_abs(pk, B, d), pk_len(pk, s), pk_len(pk, d)};
                        if (ck) F.ck_state[(size_t)(j / K - 1) * PM + row_index(en, i)] = cs;
                        if (j == j1) F.hand_state[row_index(en, i)] = cs;
                    }
                    const CkSum sum{Sm[a], slm[a], tbm[a], 0};
                    if (ck) F.ck_sum[(size_t)(j / K - 1) * C + a] = sum;
                    if (j == j1) F.hand_sum[a] = sum;
                }
            }
        }
    }

    void fill(const Job &job, const Layout &L, uint32_t track_from, Fill &F, bool allow_packed) {
        const Scoring &sc = al.opts.sc;
        const uint32_t n = job.n;
        alloc(job, L, F);
        uint32_t m_max = 0;
        for (const auto &e : L.ent) m_max = std::max(m_max, e.m);
        const uint32_t LB = (allow_packed && PACKED) ? pk_plan(sc, n, m_max) : 0;
        const uint32_t j1 = (LB && n > WINDOW + 1) ? n - WINDOW : 0;
        if (j1 > 0) { fill_packed(job, L, LB, j1, F); ++stats.launches; if (std::getenv("EMUL_TRACE")) fprintf(stderr, "PACKED n=%u j1=%u LB=%u\n", n, j1, LB); } else if (std::getenv("EMUL_TRACE")) fprintf(stderr, "WIDE n=%u LB=%u\n", n, LB);
        fill_wide(job, L, j1, std::max(track_from, j1 + 1), F);
    }

    // Re-fills contig `a` over the block of columns holding column j.
    void load_unit(const Job &job, const Layout &L, const Fill &F, uint32_t a, uint32_t j, std::vector<uint8_t> &bytes,
                   std::vector<ColRec> &ucr, TbUnit &u) {
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
        ucr.assign(je - jb, ColRec{});
        stats.cells += (uint64_t)en.m * (je - jb);
        for (uint32_t jj = jb + 1; jj <= je; ++jj) {
            const ColRec &cr = F.colrec[(size_t)jj * C + a];
            JumpInfo J{cr.jscore, cr.jlen, cr.jidx, cr.jfrom};
            ColumnArgs A{};
            A.sc = &sc; A.ent = &en; A.C = 1; A.n_tiles = en.ntiles; A.bases = al.contigs.blob.data(); A.read = job.read;
            A.j = jj; A.n = n; A.prev = st[(jj - 1) & 1].data(); A.curr = st[jj & 1].data(); A.ck_state = nullptr;
            A.tb_col = bytes.data() + (size_t)(jj - jb - 1) * pm; A.colrec_col = ucr.data() + (jj - jb - 1); A.J = &J;
            A.Sm = &Sm; A.slm = &slm; A.tbm = &tbm; A.cm = nullptr; A.cml = nullptr; A.cmk = nullptr;
            A.track = false; A.sn = nullptr; A.lastcol = false; A.last = nullptr;
            emul_column(A);
        }
        u.bytes = bytes.data(); u.cr = ucr.data(); u.a = a; u.jb = jb; u.je = je; u.pm = pm;
    }

    void run_one(const Job &job, JobResult &res) {
        const Layout &L = al.layouts.layouts[job.layout];
        const Scoring &sc = al.opts.sc;
        const uint32_t C = (uint32_t)L.ent.size(), n = job.n;
        if (n == 0) throw Error(STITCH_ERR_INVALID, "empty read");
        const bool tracked_mode = sc.ys != MIN_SCORE;   // Sn can only matter when y-suffix clipping is free
        uint32_t track_from = tracked_mode ? (n > WINDOW ? n - WINDOW + 1 : 1) : n + 1;
        Fill F;
        fill(job, L, track_from, F, true);
        if (F.need_full_track) { ++stats.launches; fill(job, L, 1, F, false); }
        for (uint32_t a = 0; a < C; ++a)
            fixup_contig(sc, L.ent[a], n, F.last.data(), F.sn.data(), tracked_mode, &F.colrec[(size_t)n * C + a].lx);
        if (const char *dump = std::getenv("STITCH_DUMP_DIR")) dump_job(dump, dump_seq++, F.last, F.sn, F.colrec, std::vector<uint8_t>());
        ReadView v;
        v.sc = sc; v.ent = L.ent.data(); v.C = C; v.n = n; v.colrec = F.colrec.data();
        v.last = F.last.data(); v.sn = F.sn.data(); v.contig_bases = al.contigs.blob.data(); v.read = job.read;
        v.pos_of = L.pos_of.data();
        v.unit.bytes = nullptr; v.unit.cr = nullptr; v.unit.a = 0xffffffffu; v.unit.jb = v.unit.je = v.unit.pm = 0;
        std::vector<uint8_t> unit_bytes; std::vector<ColRec> unit_cr;
        const uint32_t cap = 2 * n + 4 * C + 64;
        auto do_walk = [&](uint32_t a_end, RawChain &rc) -> uint32_t {
            uint32_t c = cap;
            for (;;) {
                rc.ops.assign(c, OutOp{0, 0, 0});
                WalkState ws;
                walk_begin(v, a_end, rc.ops.data(), c, ws, rc.h);
                uint32_t s;
                while ((s = walk_run(v, ws, rc.h)) == WALK_NEED_UNIT) load_unit(job, L, F, ws.a, ws.j, unit_bytes, unit_cr, v.unit);
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
