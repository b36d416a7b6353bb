// kernels_packed.cuh — the fast fill: columns [1, j1] of a read over all its contig-strands in the
// packed-key arithmetic of dp_packed.h.  Replaces the body of MCA::custom's column loop
// (multi_contig_aligner.rs:270-347 = SCA:188-239 + 292-451 + 677-697 of the reference) for the
// scorings pk_plan accepts; leaves wide checkpoints, the per-column jump records and the wide
// hand-over state from which fill_wide_kernel finishes the last columns.
//
// Structure (one CTA per read, persistent, W warps):
//   * every warp owns a contiguous chunk of 256-row tiles and walks it tile by tile; a lane owns 8
//     consecutive rows (a strip).  The rolling column state (one S key + one D key per cell) lives
//     in global memory / L2 and is updated IN PLACE: a tile is read (column j-1) and written
//     (column j) by the same warp, 128-bit accesses, 512 contiguous bytes per warp access.
//   * the diagonal and the insertion chain cross lanes with one __shfl_up each, cross tiles in
//     registers, and cross warp chunks through a 17-key halo in shared memory that the owner
//     publishes at the end of the previous column (the consumer re-derives the chain exit of the
//     strip before its chunk: the chain's reach is at most one strip, dp_packed.h).
//   * per column three CTA barriers: jump selection | tiles | per-contig finish (x-suffix tracker,
//     row m, column best incl. first-row lookup, SCA:407-429, 677-697).
#pragma once
#include "dp_packed.h"
#include "kernels_wide.cuh"

namespace stitch {
namespace gpu {

struct PackSmem {
    int32_t *Jc, *cm, *Sm, *SmKey, *tilemax, *haloS, *haloD;
    uint32_t *cml, *cmk, *slm, *tbm;
    int4 *stash;
    static size_t bytes(uint32_t cmax, uint32_t ntmax, int W) {
        return sizeof(int32_t) * ((size_t)cmax * 8 + ntmax + 2 * W * 17) + sizeof(int4) * cmax + 64;
    }
    __device__ void carve(unsigned char *raw, uint32_t cmax, uint32_t ntmax, int W) {
        stash = reinterpret_cast<int4 *>(raw);
        Jc = reinterpret_cast<int32_t *>(stash + cmax);
        cm = Jc + cmax; Sm = cm + cmax; SmKey = Sm + cmax;
        cml = reinterpret_cast<uint32_t *>(SmKey + cmax); cmk = cml + cmax; slm = cmk + cmax; tbm = slm + cmax;
        tilemax = reinterpret_cast<int32_t *>(tbm + cmax);
        haloS = tilemax + ntmax;            // [2][W][9]
        haloD = haloS + 2 * W * 9;          // [2][W][8]
    }
};

__device__ __forceinline__ uint32_t pk_sidx(uint32_t tile, uint32_t lane, uint32_t k) {
    return tile * TILE + (k >> 2) * 128u + lane * 4u + (k & 3u);
}

struct PackCtx {            // uniform per job
    PK pk; Scoring sc;
    const ContigEntry *ent; const uint16_t *owner;
    const uint8_t *bases;
    int32_t *Sst, *Dst;
    uint32_t n;
    bool yclip_mode;
};

// One tile of one column.  Returns through `prev_exit` / `prev_s7` what the next tile of the chunk needs.
template <bool SPECIAL>
__device__ __forceinline__ void pk_tile(const PackCtx &X, const PCol &pc, PackSmem &S, uint32_t tile, uint32_t lane,
                                        int32_t r0pkey, int32_t cr1key, bool chunk_start, const int32_t *hS, const int32_t *hD,
                                        int32_t &prev_exit, int32_t &prev_s7, int32_t *outS, int32_t *outD) {
    const PK &pk = X.pk;
    const uint32_t a = X.owner[tile];
    const ContigEntry en = X.ent[a];
    const uint32_t tic = tile - en.tile_start;
    const bool first = tic == 0;
    const uint32_t row0 = tic * TILE + lane * STRIP + 1;
    int32_t Sup[STRIP], Dup[STRIP];
    {
        const int4 s0 = *reinterpret_cast<const int4 *>(X.Sst + tile * TILE + lane * 4);
        const int4 s1 = *reinterpret_cast<const int4 *>(X.Sst + tile * TILE + 128 + lane * 4);
        const int4 d0 = *reinterpret_cast<const int4 *>(X.Dst + tile * TILE + lane * 4);
        const int4 d1 = *reinterpret_cast<const int4 *>(X.Dst + tile * TILE + 128 + lane * 4);
        Sup[0] = s0.x; Sup[1] = s0.y; Sup[2] = s0.z; Sup[3] = s0.w; Sup[4] = s1.x; Sup[5] = s1.y; Sup[6] = s1.z; Sup[7] = s1.w;
        Dup[0] = d0.x; Dup[1] = d0.y; Dup[2] = d0.z; Dup[3] = d0.w; Dup[4] = d1.x; Dup[5] = d1.y; Dup[6] = d1.z; Dup[7] = d1.w;
    }
    uint8_t x[STRIP];
    {
        // contig bases are 16-byte aligned per contig and a strip starts at a multiple of 8 (over-reads stay inside the blob's padding)
        const uint2 xb = *reinterpret_cast<const uint2 *>(X.bases + en.seq_off + (row0 - 1));
        x[0] = (uint8_t)xb.x; x[1] = (uint8_t)(xb.x >> 8); x[2] = (uint8_t)(xb.x >> 16); x[3] = (uint8_t)(xb.x >> 24);
        x[4] = (uint8_t)xb.y; x[5] = (uint8_t)(xb.y >> 8); x[6] = (uint8_t)(xb.y >> 16); x[7] = (uint8_t)(xb.y >> 24);
    }
    int32_t Sdg0 = __shfl_up_sync(FULL, Sup[STRIP - 1], 1);
    if (lane == 0) Sdg0 = first ? r0pkey : (chunk_start ? hS[8] : prev_s7);
    const int32_t Jc = S.Jc[a];
    PStrip st;
    int nv = STRIP; bool has_m = false;
    if (SPECIAL) {
        const int64_t left = (int64_t)en.m - (int64_t)row0;
        nv = left >= STRIP ? STRIP : (left < 0 ? 0 : (int)left);
        has_m = left >= 0 && left < STRIP;
        STITCH_UNROLL
        for (int k = 0; k < STRIP; ++k) {
            st.YC[k] = pk.NEGKEY;
            const uint32_t i = row0 + (uint32_t)k;
            if (first && X.yclip_mode && i <= en.m)
                st.YC[k] = pk_key(pk, (int64_t)X.sc.yp + X.sc.o + (int64_t)X.sc.e * i - pc.B, PP_YC, col0_slen(X.sc, i, en.m));
        }
        const bool wrap0 = first && lane == 0 && en.circular && S.tbm[a] != TB_XCLIP_SUFFIX;
        pk_pass1<true, false>(pk, pc, Sup, Dup, Sdg0, x, Jc, wrap0, pk_wbase(pk, S.SmKey[a]), nv, has_m, st);
    } else {
        pk_pass1<false, false>(pk, pc, Sup, Dup, Sdg0, x, Jc, false, 0, STRIP, false, st);
    }
    int32_t cin = pk_carry_from_exit(pk, __shfl_up_sync(FULL, st.exit, 1));
    if (first) { if (lane == 0) cin = cr1key; }
    else if (chunk_start) {
        // halo strip: rows row0-8 .. row0-1 of the same contig, from the state the previous chunk's owner published
        PStrip h;
        uint8_t hx[STRIP];
        const uint32_t hrow0 = tic * TILE - STRIP + 1;
        const uint2 xb = *reinterpret_cast<const uint2 *>(X.bases + en.seq_off + (hrow0 - 1));
        hx[0] = (uint8_t)xb.x; hx[1] = (uint8_t)(xb.x >> 8); hx[2] = (uint8_t)(xb.x >> 16); hx[3] = (uint8_t)(xb.x >> 24);
        hx[4] = (uint8_t)xb.y; hx[5] = (uint8_t)(xb.y >> 8); hx[6] = (uint8_t)(xb.y >> 16); hx[7] = (uint8_t)(xb.y >> 24);
        // halo layout: hS[0] = S of the row before the 8 halo rows, hS[1..8] = S of the halo rows, hD[0..7] = their D
        int32_t hs[STRIP], hd[STRIP];
        STITCH_UNROLL
        for (int k = 0; k < STRIP; ++k) {
            hs[k] = hS[k + 1]; hd[k] = hD[k];
            h.YC[k] = pk.NEGKEY;
            if (tic == 1 && X.yclip_mode)
                h.YC[k] = pk_key(pk, (int64_t)X.sc.yp + X.sc.o + (int64_t)X.sc.e * (hrow0 + k) - pc.B, PP_YC, col0_slen(X.sc, hrow0 + k, en.m));
        }
        pk_pass1<true, false>(pk, pc, hs, hd, hS[0], hx, Jc, false, 0, STRIP, false, h);
        if (lane == 0) cin = pk_carry_from_exit(pk, h.exit);
    } else if (lane == 0) cin = pk_carry_from_exit(pk, prev_exit);

    int32_t Sn[STRIP]; int32_t colmax = pk.NEGKEY; int32_t I_m = pk.NEGKEY; uint32_t iext_m = 0;
    if (SPECIAL) pk_pass2<true, false>(pk, pc, st, cin, 0, nv, has_m, Sn, colmax, nullptr, I_m, iext_m);
    else pk_pass2<false, false>(pk, pc, st, cin, 0, STRIP, false, Sn, colmax, nullptr, I_m, iext_m);

    // what the next tile of this chunk needs (before the in-place stores)
    prev_s7 = __shfl_sync(FULL, Sup[STRIP - 1], 31);
    prev_exit = __shfl_sync(FULL, st.exit, 31);

    if (SPECIAL) {
        STITCH_UNROLL
        for (int k = 0; k < STRIP; ++k) {
            if (k >= nv) Sn[k] = Sup[k];                       // row m is finished per contig; padding keeps its value
            if (k > nv || (k == nv && !has_m)) st.D6[k] = Dup[k];
        }
        if (has_m) S.stash[a] = make_int4(st.A[nv], st.D6[nv], st.jp[nv], I_m);
    }
    STITCH_UNROLL
    for (int k = 0; k < STRIP; ++k) { outS[k] = Sn[k]; outD[k] = st.D6[k]; }
    *reinterpret_cast<int4 *>(X.Sst + tile * TILE + lane * 4) = make_int4(Sn[0], Sn[1], Sn[2], Sn[3]);
    *reinterpret_cast<int4 *>(X.Sst + tile * TILE + 128 + lane * 4) = make_int4(Sn[4], Sn[5], Sn[6], Sn[7]);
    *reinterpret_cast<int4 *>(X.Dst + tile * TILE + lane * 4) = make_int4(st.D6[0], st.D6[1], st.D6[2], st.D6[3]);
    *reinterpret_cast<int4 *>(X.Dst + tile * TILE + 128 + lane * 4) = make_int4(st.D6[4], st.D6[5], st.D6[6], st.D6[7]);
    STITCH_UNROLL
    for (int d = 16; d >= 1; d >>= 1) colmax = pk_max(colmax, __shfl_xor_sync(FULL, colmax, d));
    if (lane == 0) S.tilemax[tile] = colmax;
}


template <int W>
__global__ void __launch_bounds__(W * 32) fill_packed_kernel(const Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PackSmem S; S.carve(smem_raw, P.cmax, P.ntmax, W);
    __shared__ uint32_t sJob;
    __shared__ PCol s_pc[2];
    __shared__ int32_t s_r0pkey[2], s_cr1key[2];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr uint32_t T = W * 32;
    const Scoring sc = P.sc;

    for (;;) {
        __syncthreads();
        if (tid == 0) sJob = atomicAdd(P.counter, 1u);
        __syncthreads();
        if (sJob >= P.n_jobs) break;
        const JobDesc jd = P.jobs[P.order[sJob]];
        const LayoutDesc ld = P.layouts[jd.layout];
        const ContigEntry *ent = P.ents + ld.ent_off;
        const uint16_t *owner = P.owners + ld.owner_off;
        const uint32_t C = ld.C, NT = ld.n_tiles, PM = ld.PM, n = jd.n, K = P.K, j1 = jd.j0;
        const PK pk = pk_make(sc, jd.LB);
        PackCtx X;
        X.pk = pk; X.sc = sc; X.ent = ent; X.owner = owner; X.bases = P.contig_bases;
        X.Sst = P.pstate + (uint64_t)blockIdx.x * P.pstate_stride; X.Dst = X.Sst + P.pstate_half;
        X.n = n; X.yclip_mode = sc.yp != MIN_SCORE && sc.xp == MIN_SCORE;
        ColRec *colrec = P.colrec + jd.colrec_off;
        int32_t *gcol = P.gcol + jd.gcol_off;
        const uint8_t *read = P.reads + jd.read_off;
        const uint32_t Weff = NT < (uint32_t)W ? NT : (uint32_t)W;

        // ---- column 0 (SCA:97-186), base B_0 = 0 ----
        for (uint32_t p = tid; p < PM; p += T) { X.Sst[p] = pk.NEGKEY; X.Dst[p] = pk.NEGKEY + pk.PD6; }
        __syncthreads();
        for (uint32_t tile = warp; tile < NT; tile += W) {
            const ContigEntry en = ent[owner[tile]];
            const uint32_t tic = tile - en.tile_start;
            STITCH_UNROLL
            for (int k = 0; k < STRIP; ++k) {
                const uint32_t i = tic * TILE + lane * STRIP + (uint32_t)k + 1;
                if (i <= en.m) {
                    const Col0 c0 = col0_at(sc, i, en.m);
                    X.Sst[pk_sidx(tile, lane, (uint32_t)k)] = pk_from_wide(pk, 0, c0.S, c0.sl, 0);
                }
            }
        }
        for (uint32_t a = tid; a < C; a += T) {
            const ContigEntry en = ent[a];
            S.cm[a] = 0; S.cml[a] = 0; S.cmk[a] = 0;
            const Col0 cmm = col0_at(sc, en.m, en.m);
            S.Sm[a] = cmm.S; S.slm[a] = cmm.sl; S.tbm[a] = cmm.s_tb;
            S.SmKey[a] = pk_from_wide(pk, 0, cmm.S, cmm.sl, 0);
        }
        if (tid == 0) { s_pc[0].B = 0; s_pc[0].delta = 0; }
        __syncthreads();
        // halos of column 0 (read by column 1 from parity slot 0): the 9 rows before every chunk
        if (warp >= 1 && warp < Weff && lane < 9) {
            const uint32_t t_lo = (uint32_t)((uint64_t)NT * warp / Weff);
            const uint32_t hl = lane == 0 ? 30u : 31u, hk = lane == 0 ? (uint32_t)STRIP - 1 : lane - 1;
            const uint32_t pi = pk_sidx(t_lo - 1, hl, hk);
            S.haloS[(0 * W + warp) * 9 + lane] = X.Sst[pi];
            if (lane >= 1) S.haloD[(0 * W + warp) * 8 + lane - 1] = X.Dst[pi];
        }

        for (uint32_t j = 1; j <= j1 + 1; ++j) {
            const uint32_t par = j & 1u;
            // ---- phase S: base of the column, jump selection (MCA:279-331), per-column constants ----
            {
                const int32_t Bprev = s_pc[par ^ 1u].B;
                if (tid < C || tid == 0) {
                    int32_t g = S.cm[0];
                    for (uint32_t a = 1; a < C; ++a) g = S.cm[a] > g ? S.cm[a] : g;
                    PCol pcl; pcl.B = g; pcl.delta = g - Bprev;
                    for (uint32_t a = tid; a < C; a += T) {
                        const JumpInfo J = select_jump(sc, ent, C, a, S.cm, S.cml, S.cmk);
                        ColRec cr; cr.jscore = J.score; cr.jlen = J.len; cr.jidx = J.idx; cr.jfrom = J.from;
                        cr.lx = 0; cr.pad0 = cr.pad1 = cr.pad2 = 0;
                        colrec[(uint64_t)j * C + a] = cr;
                        S.Jc[a] = pk_jc(pk, pcl, J.score, J.len);
                    }
                    if (tid == 0) {
                        gcol[j - 1] = g;
                        if (j <= j1) {
                            s_pc[par] = pk_col(pk, sc, g, Bprev, j, n, read[j - 1]);
                            const Row0 r0 = row0_at(sc, j, n), r0p = row0_at(sc, j - 1, n);
                            s_r0pkey[par] = pk_from_wide(pk, Bprev, r0p.S, r0p.sl, 0);
                            s_cr1key[par] = pk_carry_row1(pk, s_pc[par], sc, r0);
                        }
                    }
                }
            }
            if (j == j1 + 1) break;   // only the jump records of the first wide column were needed
            __syncthreads();
            const PCol pc = s_pc[par];
            const int32_t r0pkey = s_r0pkey[par], cr1key = s_cr1key[par];

            // ---- phase T: tiles ----
            if (warp < Weff) {
                const uint32_t t_lo = (uint32_t)((uint64_t)NT * warp / Weff), t_hi = (uint32_t)((uint64_t)NT * (warp + 1) / Weff);
                const int32_t *hS = S.haloS + ((par ^ 1u) * W + warp) * 9, *hD = S.haloD + ((par ^ 1u) * W + warp) * 8;
                int32_t prev_exit = 0, prev_s7 = 0;
                int32_t oS[STRIP], oD[STRIP];
                for (uint32_t tile = t_lo; tile < t_hi; ++tile) {
                    const ContigEntry en = ent[owner[tile]];
                    const uint32_t tic = tile - en.tile_start;
                    if (tic == 0 || tic + 1 == en.ntiles)
                        pk_tile<true>(X, pc, S, tile, lane, r0pkey, cr1key, tile == t_lo, hS, hD, prev_exit, prev_s7, oS, oD);
                    else
                        pk_tile<false>(X, pc, S, tile, lane, r0pkey, cr1key, tile == t_lo, hS, hD, prev_exit, prev_s7, oS, oD);
                }
                if (warp + 1 < Weff) {   // publish the halo of the next chunk for the next column
                    int32_t *nS = S.haloS + (par * W + warp + 1) * 9, *nD = S.haloD + (par * W + warp + 1) * 8;
                    if (lane == 31) {
                        STITCH_UNROLL
                        for (int k = 0; k < STRIP; ++k) { nS[k + 1] = oS[k]; nD[k] = oD[k]; }
                    }
                    if (lane == 30) nS[0] = oS[STRIP - 1];
                }
            }
            __syncthreads();

            // ---- phase F: per contig, tracker + row m + column best ----
            const Row0 r0 = row0_at(sc, j, n);
            for (uint32_t a = warp; a < C; a += W) {
                const ContigEntry en = ent[a];
                int32_t kmax = pk.NEGKEY;
                for (uint32_t t = lane; t < en.ntiles; t += 32) kmax = pk_max(kmax, S.tilemax[en.tile_start + t]);
                STITCH_UNROLL
                for (int d = 16; d >= 1; d >>= 1) kmax = pk_max(kmax, __shfl_xor_sync(FULL, kmax, d));
                const int32_t smax = pk_rel(pk, kmax);
                uint32_t frow = 0xffffffffu; int32_t fkey = 0;
                if (en.m >= 2) {
                    uint32_t ft = 0xffffffffu;
                    for (uint32_t t = lane; t < en.ntiles; t += 32)
                        if (pk_rel(pk, S.tilemax[en.tile_start + t]) == smax) { ft = t; break; }
                    STITCH_UNROLL
                    for (int d = 16; d >= 1; d >>= 1) { const uint32_t o = __shfl_xor_sync(FULL, ft, d); ft = o < ft ? o : ft; }
                    const uint32_t tile = en.tile_start + ft;
                    const int4 s0 = *reinterpret_cast<const int4 *>(X.Sst + tile * TILE + lane * 4);
                    const int4 s1 = *reinterpret_cast<const int4 *>(X.Sst + tile * TILE + 128 + lane * 4);
                    const int32_t sk[STRIP] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                    STITCH_UNROLL
                    for (int k = STRIP - 1; k >= 0; --k) {
                        const uint32_t i = ft * TILE + lane * STRIP + (uint32_t)k + 1;
                        if (i < en.m && pk_rel(pk, sk[k]) == smax) { frow = i; fkey = sk[k]; }
                    }
                    uint32_t best = frow;
                    STITCH_UNROLL
                    for (int d = 16; d >= 1; d >>= 1) { const uint32_t o = __shfl_xor_sync(FULL, best, d); best = o < best ? o : best; }
                    const uint32_t src = __ffs(__ballot_sync(FULL, frow == best)) - 1;
                    fkey = __shfl_sync(FULL, fkey, src);
                    frow = best;
                }
                if (lane == 0) {
                    CmPart rows; cm_init(rows);
                    XsPart tr; xs_init(tr);
                    if (en.m >= 2) {
                        rows.S = pc.B + smax; rows.row = frow; rows.sl = pk_len(pk, fkey); rows.valid = 1;
                        if (sc.xs != MIN_SCORE) { tr.t = pc.B + smax + sc.xs; tr.len = pk_len(pk, kmax); tr.row = 1; }
                    }
                    const int4 sm_ = S.stash[a];
                    RowM rm;
                    rm.diag = pk_abs(pk, pc.B, sm_.x); rm.dgl = pk_len(pk, sm_.x);
                    rm.D = pk_abs(pk, pc.B, sm_.y); rm.dl = pk_len(pk, sm_.y); rm.dext = 0;
                    rm.I = pk_abs(pk, pc.B, sm_.w); rm.il = pk_len(pk, sm_.w); rm.iext = 0;
                    rm.jp.score = pk_abs(pk, pc.B, sm_.z); rm.jp.len = pk_len(pk, sm_.z); rm.jp.idx = 0; rm.jp.from = 0;
                    { const int32_t dj = sc.o + sc.e * (int32_t)j; rm.xclip = sc.xp + (sc.yp > dj ? sc.yp : dj); }
                    rm.xclip_len = r0.sl;
                    rm.yclip = sc.yp + sc.o + sc.e * (int32_t)en.m; rm.yclip_len = 0;
                    rm.is_match = P.contig_bases[en.seq_off + en.m - 1] == pc.q;
                    const RowMOut ro = finish_rowm(sc, rm, tr, en.contig_idx, en.m);
                    const uint32_t r = en.m - 1;
                    const int32_t skey = pk_from_wide(pk, pc.B, ro.c.S, ro.c.sl, 0);
                    X.Sst[pk_sidx(en.tile_start + r / TILE, (r % TILE) / STRIP, r % STRIP)] = skey;
                    CmPart cmv; cm_init(cmv);
                    cm_add(cmv, r0.S, r0.sl, 0);
                    cmv = cm_merge(cmv, rows);
                    CmPart top; top.S = ro.c.S; top.row = en.m; top.sl = ro.c.sl; top.valid = 1;
                    cmv = cm_merge(cmv, top);
                    S.cm[a] = cmv.S; S.cmk[a] = cmv.row; S.cml[a] = cmv.sl;
                    S.Sm[a] = ro.c.S; S.slm[a] = ro.c.sl; S.tbm[a] = ro.s_tb; S.SmKey[a] = skey;
                }
            }
            __syncthreads();

            // ---- wide checkpoints (every K columns) and the hand-over state at column j1 ----
            const bool ck = (j % K == 0) && j < n;
            if (ck || j == j1) {
                CellState *dck = ck ? P.ck_state + jd.ck_off + (uint64_t)(j / K - 1) * PM : nullptr;
                CellState *dh = (j == j1) ? P.hand_state + jd.hand_off : nullptr;
                for (uint32_t idx = tid; idx < PM; idx += T) {
                    const uint32_t tile = idx / TILE, w = idx % TILE, k = w / 32, ln = w % 32;
                    const uint32_t pi = pk_sidx(tile, ln, k);
                    const int32_t s = X.Sst[pi], d = X.Dst[pi];
                    CellState cs; cs.S = pk_abs(pk, pc.B, s); cs.D = pk_abs(pk, pc.B, d); cs.sl = pk_len(pk, s); cs.dl = pk_len(pk, d);
                    if (dck) dck[idx] = cs;
                    if (dh) dh[idx] = cs;
                }
                for (uint32_t a = tid; a < C; a += T) {
                    CkSum cs; cs.Sm = S.Sm[a]; cs.slm = S.slm[a]; cs.tbm = S.tbm[a]; cs.pad = 0;
                    if (ck) P.ck_sum[jd.cksum_off + (uint64_t)(j / K - 1) * C + a] = cs;
                    if (j == j1) P.hand_sum[jd.handsum_off + a] = cs;
                }
            }
        }
    }
}

}  // namespace gpu
}  // namespace stitch
