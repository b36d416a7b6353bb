// kernels_packed.cuh — the fast path: the DP columns of a read in the packed-key arithmetic of
// dp_packed.h.  Replaces the body of MCA::custom's column loop (multi_contig_aligner.rs:270-347 =
// SCA:188-239 + 292-451 + 677-697 of the reference) for the scorings pk_plan accepts.
//
//   fill_packed_kernel   bulk: all n columns over all contig-strands, no traceback output; leaves
//                        column-state checkpoints every K columns, the jump record of every
//                        (contig, column), the best score of every column, and the column the tail
//                        must restart from.
//   pk_tail (same kernel) the last columns again, from a checkpoint, in the traceback variant: keeps
//                        the y-suffix trackers of every row (SCA:432-447) for every column that can
//                        still hold a final tracker value, and the column-n records the end-of-read
//                        fix-up edits (SCA:453-555).
//   pk_refill_unit       (walk kernel) packed traceback bytes of one contig over one block of
//                        columns, re-filled from a checkpoint (checkpoint-and-recompute).
//
// Structure of a column (one CTA per read, W warps):
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
#include <cooperative_groups.h>
#include <cuda_pipeline_primitives.h>

#include "dp_packed.h"
#include "kernels_wide.cuh"

namespace stitch {
namespace gpu {

namespace cg = cooperative_groups;

// Bulk asynchronous copies (cp.async.bulk, the TMA engine without a tensor map: SASS UBLKCP) completing on an mbarrier
// (SYNCS).  One elected lane issues a whole tile; the copy engine works beside the LSU pipe, no per-lane address arithmetic.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
// Orders this thread's earlier generic-proxy writes (st.global of the rolling state) before later async-proxy reads of the
// same bytes (the bulk copy of the tile in the next column); executed by every writer before the CTA barrier.
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async;" ::: "memory"); }

// The CTAs that share one read: a thread-block cluster (size > 1) or a single CTA.  A cluster splits the tiles
// of every column over its CTAs; the small per-contig tables live in every CTA's shared memory and are kept
// identical through distributed-shared-memory stores, ordered by two cluster barriers per column.
struct Team {
    uint32_t rank, size;
    __device__ __forceinline__ void sync() const { if (size > 1) cg::this_cluster().sync(); else __syncthreads(); }
    template <typename T>
    __device__ __forceinline__ T *peer(T *p, uint32_t r) const { return size > 1 ? cg::this_cluster().map_shared_rank(p, r) : p; }
};

struct PackSmem {
    int32_t *Jc, *cm, *Sm, *SmKey, *DmKey, *tilemax, *haloS, *haloD;   // DmKey: D key of row m of every contig (latest column)
    uint32_t *cml, *cmk, *slm, *tbm, *haloF;
    PkRowM *stash;
    JumpInfo *Jw;
    PkQuiet *Q;             // [2][cmax]: closed form of a quiet tile per contig, by column parity (dp_packed.h)
    // per tile, one byte: bits 0..4 base classes present (A C G T other), bit 5 quiet flag of the latest column, bit 6 the
    // tile may turn quiet at all (static: not the last tile of its contig, or last tiles may); owned by the tile's warp
    uint8_t *tb;
    static constexpr uint32_t TB_MASK = 31u, TB_Q = 32u, TB_CANQ = 64u;
    ContigEntry *ent_s;     // [cmax]: the layout's contig table (bulk pass: no global loads in the per-column serial phases)
    uint16_t *owner_s;      // [ntmax]: contig position of every tile
    uint16_t *clist;        // [ntmax]: per warp chunk (slice [t_lo, t_hi)), the chunk-relative offsets of the tiles computed in this column
    uint32_t cmax;
    unsigned char *stage;   // [W][depth][STAGE_BYTES]: bulk-copy ring of the next tiles (state in global memory only)
    unsigned long long *mbar;   // [W][4]: one mbarrier per ring slot
    uint32_t *mphase;       // [W]: phase parity of every slot's mbarrier (bit = slot), kept across columns and reads
    // one slot: S keys, D keys, the 16 bases before the tile (its upper neighbour's last rows: diagonal source of a skipped
    // neighbour, halo strip of a chunk's first tile), the tile's bases
    static constexpr uint32_t STAGE_PRE = 16, STAGE_BASES = 2 * TILE * 4 + STAGE_PRE;
    static constexpr uint32_t STAGE_BYTES = 2 * TILE * 4 + STAGE_PRE + TILE;
    // `stage_bytes`: the front area = bulk-copy ring of the tiles (default_stage), or the cluster's slice of the rolling
    // state when that lives in shared memory, or the walk phase's re-fill state; 0 = none (walk kernel)
    static size_t default_stage(int W, uint32_t depth = 2) { return (size_t)W * depth * STAGE_BYTES; }
    // One definition of the layout for the host (sizes, Params::pso) and the device (carve).
    static __host__ __device__ PackSmemOff layout(uint32_t cmax, uint32_t ntmax, int W, size_t stage_bytes) {
        PackSmemOff o;
        uint32_t p = (uint32_t)stage_bytes;
        o.Jw = p; p += (uint32_t)sizeof(JumpInfo) * cmax;
        o.stash = p; p += (uint32_t)sizeof(PkRowM) * cmax;
        o.Jc = p; p += 4u * cmax; o.cm = p; p += 4u * cmax; o.Sm = p; p += 4u * cmax; o.SmKey = p; p += 4u * cmax;
        o.cml = p; p += 4u * cmax; o.cmk = p; p += 4u * cmax; o.slm = p; p += 4u * cmax; o.tbm = p; p += 4u * cmax;
        o.DmKey = p; p += 4u * cmax;
        o.tilemax = p; p += 4u * ntmax;
        o.haloS = p; p += 4u * 2u * (uint32_t)W * 9u;            // [2][W][9]
        o.haloD = p; p += 4u * 2u * (uint32_t)W * 8u;            // [2][W][8]
        o.haloF = p; p += 4u * (2u * (uint32_t)W + 2u);          // [2][W]: quiet flag of the last tile of the previous chunk
        o.Q = p; p += (uint32_t)sizeof(PkQuiet) * 2u * cmax;
        o.ent_s = p; p += (uint32_t)sizeof(ContigEntry) * cmax;
        o.owner_s = p; p += 2u * ntmax;
        o.clist = p; p += 2u * ntmax;
        o.tb = p; p += ntmax;
        p = (p + 7u) & ~7u;
        o.mbar = p; p += 8u * 4u * (uint32_t)W;
        o.mphase = p; p += 4u * (uint32_t)W;
        o.end = p;
        return o;
    }
    static size_t bytes(uint32_t cmax, uint32_t ntmax, int W, size_t stage_bytes) {
        return stage_bytes + sizeof(int32_t) * ((size_t)cmax * 9 + ntmax + 2 * W * 18) +
               (sizeof(PkRowM) + sizeof(JumpInfo) + 2 * sizeof(PkQuiet) + sizeof(ContigEntry)) * cmax + 5 * (size_t)ntmax + 4 * 8 + 64 + 36 * (size_t)W + 8;
    }
    __device__ __forceinline__ void carve(unsigned char *raw, const PackSmemOff &o, uint32_t cmax_) {
        cmax = cmax_;
        stage = raw;
        Jw = reinterpret_cast<JumpInfo *>(raw + o.Jw);
        stash = reinterpret_cast<PkRowM *>(raw + o.stash);
        Jc = reinterpret_cast<int32_t *>(raw + o.Jc); cm = reinterpret_cast<int32_t *>(raw + o.cm);
        Sm = reinterpret_cast<int32_t *>(raw + o.Sm); SmKey = reinterpret_cast<int32_t *>(raw + o.SmKey);
        cml = reinterpret_cast<uint32_t *>(raw + o.cml); cmk = reinterpret_cast<uint32_t *>(raw + o.cmk);
        slm = reinterpret_cast<uint32_t *>(raw + o.slm); tbm = reinterpret_cast<uint32_t *>(raw + o.tbm);
        DmKey = reinterpret_cast<int32_t *>(raw + o.DmKey);
        tilemax = reinterpret_cast<int32_t *>(raw + o.tilemax);
        haloS = reinterpret_cast<int32_t *>(raw + o.haloS); haloD = reinterpret_cast<int32_t *>(raw + o.haloD);
        haloF = reinterpret_cast<uint32_t *>(raw + o.haloF);
        Q = reinterpret_cast<PkQuiet *>(raw + o.Q);
        ent_s = reinterpret_cast<ContigEntry *>(raw + o.ent_s);
        owner_s = reinterpret_cast<uint16_t *>(raw + o.owner_s);
        clist = reinterpret_cast<uint16_t *>(raw + o.clist);
        tb = raw + o.tb;
        mbar = reinterpret_cast<unsigned long long *>(raw + o.mbar);
        mphase = reinterpret_cast<uint32_t *>(raw + o.mphase);
    }
};

// State layout: per tile 2 KB contiguous, [256 S keys][256 D keys], each as [k / 4][lane][k % 4] so that a lane's 8 keys
// are two 16-byte words and a warp access is 512 contiguous bytes.  Dst = Sst + TILE.
constexpr uint32_t ST = 2 * TILE;
__device__ __forceinline__ uint32_t pk_sidx(uint32_t tile, uint32_t lane, uint32_t k) {
    return tile * ST + (k >> 2) * 128u + lane * 4u + (k & 3u);
}

struct PackCtx {            // uniform per (job, set of contigs)
    PK pk; Scoring sc;
    const ContigEntry *ent; const uint16_t *owner;   // owner == nullptr: a single contig (position 0)
    uint32_t C, NT;
    const uint8_t *bases;
    const uint8_t *tbases;  // staged fills: the layout's bases in tile order (tile t at tbases + t * TILE, 16 readable bytes before tile 0)
    int32_t *Sst, *Dst;
    uint32_t n;
    bool yclip_mode;
    bool state_smem;        // the state arrays live in shared memory (walk kernel re-fills)
    bool staged;            // tiles staged by bulk copies (state and bases both in global memory, stage buffers carved)
    uint32_t stage_depth;   // slots per warp of the staging ring (2 or 4): depth - 1 tiles are in flight ahead of the one computed
    Team team;
    // Tiles [own_lo, own_hi) are this CTA's (all tiles for a single-CTA team).  Sst/Dst address them as
    // Sst + tile * ST.  With cluster_smem the state of a tile lives in the shared memory of the CTA that owns it
    // (cstate = this CTA's block, same offset in every CTA of the cluster) and Sst = cstate - own_lo * ST.
    uint32_t own_lo, own_hi, warps;
    // Cone re-fills of the walk (TbUnit, dp_core.h) run the columns over a WINDOW of the contig's tiles: tiles
    // [win_lo, win_lo + NT), one fixed chunk per warp; chunks entirely below tile `skip_below` are left stale in this column; no per-contig finish
    // (row m is outside the window).  Everything else: win_lo = 0, skip_below = 0, no_finish = false.
    uint32_t win_lo, skip_below;
    bool no_finish;
    bool cluster_smem;
    bool quiet;             // the bulk pass may skip quiet tiles (single-CTA teams with the state in global memory)
    bool quiet_first, quiet_edge, quiet_last;   // the first / last tile of a contig, the first and last tile of a warp chunk may be skipped too
    int32_t *cstate;
    const uint32_t *cta_lo;   // shared memory: first tile of every CTA of the team, [size + 1]
};

// First tile of the chunk of global warp `gw` (chunks are contiguous and cover [0, NT)).
__device__ __forceinline__ uint32_t pk_chunk_lo(uint32_t NT, uint32_t Weff, uint32_t gw) {
    return gw >= Weff ? NT : (uint32_t)((uint64_t)NT * gw / Weff);
}
__device__ __forceinline__ void pk_set_ownership(PackCtx &X, uint32_t W) {
    const uint32_t GW = X.team.size * W, Weff = X.NT < GW ? X.NT : GW;
    X.warps = W;
    X.own_lo = X.win_lo + pk_chunk_lo(X.NT, Weff, X.team.rank * W);
    X.own_hi = X.win_lo + pk_chunk_lo(X.NT, Weff, (X.team.rank + 1) * W);
}
// The 512-key block of any tile of the read (possibly in another CTA's shared memory).
__device__ __forceinline__ int32_t *pk_tile_ptr(const PackCtx &X, uint32_t tile) {
    if (!X.cluster_smem) return X.Sst + tile * ST;
    uint32_t r = 0;
    while (r + 1 < X.team.size && X.cta_lo[r + 1] <= tile) ++r;
    return X.team.peer(X.cstate, r) + (tile - X.cta_lo[r]) * ST;
}

// State loads: through the L2 only when the state is in global memory (it streams, and row m is rewritten by
// another CTA of the team), plain when it lives in shared memory.
__device__ __forceinline__ int4 pk_ld_state(const PackCtx &X, const int32_t *p) {
    return X.state_smem ? *reinterpret_cast<const int4 *>(p) : __ldcg(reinterpret_cast<const int4 *>(p));
}

struct PkColOut {           // traceback-variant outputs of a column (all optional)
    uint8_t *tb_col;        // packed traceback bytes, linear rows
    ColRec *colrec_col;     // per contig: Lx[j] (+ the jump record)
    SnRec *sn; LastCell *last;
    bool track, lastcol;
    int32_t track_thr;      // only cells with S >= track_thr can hold a final tracker value (absolute score)
};

__device__ __forceinline__ void unpack8(const uint2 xb, uint8_t *x) {
    x[0] = (uint8_t)xb.x; x[1] = (uint8_t)(xb.x >> 8); x[2] = (uint8_t)(xb.x >> 16); x[3] = (uint8_t)(xb.x >> 24);
    x[4] = (uint8_t)xb.y; x[5] = (uint8_t)(xb.y >> 8); x[6] = (uint8_t)(xb.y >> 16); x[7] = (uint8_t)(xb.y >> 24);
}

// One tile of one column.  `prev_*` carry what the next tile of the same chunk needs.
// Quiet tiles (QUIET, bulk pass only; dp_packed.h): `qz.mat` = the tile's own state of column j-1 is the closed form
// (its memory may be stale): materialise it instead of loading; `qz.prev_skipped` = the previous tile of this chunk
// was skipped in this column (its last row and its chain exit come from the closed form).  Returns whether every
// ordinary cell of the tile is in the closed form of column j.
struct PkColStat { uint32_t skipped, timing; unsigned long long t_tiles, t_finish, t_busy, t_select, t_f1, t_f2, t_fa, t_warp[32], n_warp[32]; };   // per job, shared memory
struct PkQuietArgs {
    bool mat, prev_skipped;
    const PkQuiet *Qp, *Qn;     // closed forms of column j-1 / j of this tile's contig (shared memory)
    const uint8_t *yq;          // yq[k] = read base y_{j-k}, 0 when there is none
    int32_t deadrel;
};
__device__ __forceinline__ int32_t pk_quiet_D_dev(const PK &pk, const PkQuiet &q, uint8_t xb, const uint8_t *yq) {
    int32_t d = xb == yq[1] ? q.t[0][0] : q.t[0][1];
    STITCH_UNROLL
    for (int k = 1; k < PKQ_L; ++k) d = pk_max(d, xb == yq[k + 1] ? q.t[k][0] : q.t[k][1]);
    return pk_clean6(pk, d);
}

template <bool SPECIAL, bool TB, bool QUIET>
__device__ __forceinline__ bool pk_tile(const PackCtx &X, const PCol &pc, PackSmem &S, const PkColOut &O, uint32_t j, uint32_t tile,
                                        uint32_t lane, int32_t r0pkey, int32_t cr1key, bool chunk_start, const int32_t *hS,
                                        const int32_t *hD, int32_t &prev_exit, uint32_t &prev_exit_open, int32_t &prev_s7,
                                        const unsigned char *stg, uint32_t a, const ContigEntry &en, int32_t Jc, const PkQuietArgs &qz) {
    const PK &pk = X.pk;
    const uint32_t tic = tile - en.tile_start;
    const bool first = tic == 0;
    const uint32_t row0 = tic * TILE + lane * STRIP + 1;
    int32_t Sup[STRIP], Dup[STRIP];
    uint2 xb;
    if (QUIET && qz.mat) {
        xb = stg ? reinterpret_cast<const uint2 *>(stg + PackSmem::STAGE_BASES)[lane] : *reinterpret_cast<const uint2 *>(X.bases + en.seq_off + (row0 - 1));
        uint8_t xm[STRIP];
        unpack8(xb, xm);
        const PkQuiet qp = *qz.Qp;
        STITCH_UNROLL
        for (int k = 0; k < STRIP; ++k) {
            Sup[k] = xm[k] == qz.yq[1] ? qp.bk[0] : qp.bk[1];
            Dup[k] = pk_quiet_D_dev(pk, qp, xm[k], qz.yq + 1);
            // row m is never in the closed form: its keys of column j-1 are kept per contig
            if (SPECIAL && row0 + (uint32_t)k == en.m) { Sup[k] = S.SmKey[a]; Dup[k] = S.DmKey[a]; }
        }
    } else {
        int4 s0, s1, d0, d1;
        if (stg) {   // this tile was staged in shared memory by a bulk copy while the previous tile was computed
            const int4 *q = reinterpret_cast<const int4 *>(stg);
            s0 = q[lane]; s1 = q[32 + lane]; d0 = q[64 + lane]; d1 = q[96 + lane];
            xb = reinterpret_cast<const uint2 *>(stg + PackSmem::STAGE_BASES)[lane];
        } else {
            s0 = pk_ld_state(X, X.Sst + tile * ST + lane * 4);
            s1 = pk_ld_state(X, X.Sst + tile * ST + 128 + lane * 4);
            d0 = pk_ld_state(X, X.Dst + tile * ST + lane * 4);
            d1 = pk_ld_state(X, X.Dst + tile * ST + 128 + lane * 4);
            // contig bases are 16-byte aligned per contig and a strip starts at a multiple of 8 (over-reads stay inside the blob's padding)
            xb = *reinterpret_cast<const uint2 *>(X.bases + en.seq_off + (row0 - 1));
        }
        Sup[0] = s0.x; Sup[1] = s0.y; Sup[2] = s0.z; Sup[3] = s0.w; Sup[4] = s1.x; Sup[5] = s1.y; Sup[6] = s1.z; Sup[7] = s1.w;
        Dup[0] = d0.x; Dup[1] = d0.y; Dup[2] = d0.z; Dup[3] = d0.w; Dup[4] = d1.x; Dup[5] = d1.y; Dup[6] = d1.z; Dup[7] = d1.w;
    }
    uint8_t x[STRIP];
    unpack8(xb, x);
    int32_t Sdg0 = __shfl_up_sync(FULL, Sup[STRIP - 1], 1);
    if (lane == 0) {
        if (first) Sdg0 = r0pkey;
        else if (chunk_start) Sdg0 = hS[8];
        else if (QUIET && qz.prev_skipped)   // last row of a quiet tile (its base: the byte before this tile's bases)
            Sdg0 = (stg ? stg[PackSmem::STAGE_BASES - 1] : X.bases[en.seq_off + row0 - 2]) == qz.yq[1] ? qz.Qp->bk[0] : qz.Qp->bk[1];
        else Sdg0 = prev_s7;
    }
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
        pk_pass1<true, TB>(pk, pc, Sup, Dup, Sdg0, x, Jc, wrap0, pk_wbase(pk, S.SmKey[a]), nv, has_m, st);
    } else {
        pk_pass1<false, TB>(pk, pc, Sup, Dup, Sdg0, x, Jc, false, 0, STRIP, false, st);
    }
    int32_t cin = pk_carry_from_exit(pk, __shfl_up_sync(FULL, st.exit, 1));
    uint32_t cin_open = 0;
    if (TB) cin_open = __shfl_up_sync(FULL, st.exit_open, 1);
    if (first) { if (lane == 0) { cin = cr1key; cin_open = 1; } }
    else if (chunk_start) {
        // halo strip: rows row0-8 .. row0-1 of the same contig, from the state the previous chunk's owner published
        // (hS[0] = S of the row before the 8 halo rows, hS[1..8] = S of the halo rows, hD[0..7] = their D)
        PStrip h;
        uint8_t hx[STRIP];
        const uint32_t hrow0 = tic * TILE - STRIP + 1;
        unpack8(stg ? *reinterpret_cast<const uint2 *>(stg + PackSmem::STAGE_BASES - STRIP) : *reinterpret_cast<const uint2 *>(X.bases + en.seq_off + (hrow0 - 1)), hx);
        int32_t hs[STRIP], hd[STRIP];
        STITCH_UNROLL
        for (int k = 0; k < STRIP; ++k) {
            hs[k] = hS[k + 1]; hd[k] = hD[k];
            h.YC[k] = pk.NEGKEY;
            if (tic == 1 && X.yclip_mode)
                h.YC[k] = pk_key(pk, (int64_t)X.sc.yp + X.sc.o + (int64_t)X.sc.e * (hrow0 + k) - pc.B, PP_YC, col0_slen(X.sc, hrow0 + k, en.m));
        }
        pk_pass1<true, TB>(pk, pc, hs, hd, hS[0], hx, Jc, false, 0, STRIP, false, h);
        if (lane == 0) { cin = pk_carry_from_exit(pk, h.exit); if (TB) cin_open = h.exit_open; }
    } else if (lane == 0) {
        if (QUIET && qz.prev_skipped) { cin = pk.NEGKEY + pk.PI5; cin_open = 0; }   // the chain out of a quiet tile is dead (C4)
        else { cin = pk_carry_from_exit(pk, prev_exit); cin_open = prev_exit_open; }
    }

    int32_t Sn[STRIP], Iarr[STRIP]; uint8_t tbb[STRIP];
    int32_t colmax = pk.NEGKEY; int32_t I_m = pk.NEGKEY; uint32_t iext_m = 0;
    int32_t notq = 0;   // ordinary tiles of the bulk pass: the S half of the closed-form test rides along in pass 2
    if (SPECIAL) pk_pass2<true, TB>(pk, pc, st, cin, cin_open, nv, has_m, Sn, colmax, tbb, Iarr, I_m, iext_m);
    else pk_pass2<false, TB>(pk, pc, st, cin, cin_open, STRIP, false, Sn, colmax, tbb, Iarr, I_m, iext_m, (QUIET && !TB) ? &notq : nullptr);

    // what the next tile of this chunk needs (before the in-place stores)
    prev_s7 = __shfl_sync(FULL, Sup[STRIP - 1], 31);
    prev_exit = __shfl_sync(FULL, st.exit, 31);
    if (TB) prev_exit_open = __shfl_sync(FULL, st.exit_open, 31);

    if (TB) {   // traceback bytes, y-suffix trackers, column-n records of the ordinary rows
        if (O.tb_col) {
            uint32_t lo = 0, hi = 0;
            STITCH_UNROLL
            for (int k = 0; k < 4; ++k) {
                lo |= (uint32_t)((!SPECIAL || k < nv) ? tbb[k] : 0) << (8 * k);
                hi |= (uint32_t)((!SPECIAL || k + 4 < nv) ? tbb[k + 4] : 0) << (8 * k);
            }
            *reinterpret_cast<uint2 *>(O.tb_col + tile * TILE + lane * STRIP) = make_uint2(lo, hi);
        }
        if (O.track || O.lastcol) {
            const int32_t thrkey = pk_key(pk, (int64_t)O.track_thr - pc.B, 0, 0);
            STITCH_UNROLL
            for (int k = 0; k < STRIP; ++k) {
                const bool trk = O.track && Sn[k] >= thrkey;
                if ((!SPECIAL || k < nv) && (trk || O.lastcol)) {
                    const JumpInfo Jw = S.Jw[a];
                    const uint32_t si = state_index(tile, lane, (uint32_t)k);
                    pk_cell_records(pk, pc, X.sc, Sn[k], Iarr[k], tbb[k], x[k] == pc.q, en.contig_idx, row0 + (uint32_t)k, en.m, Jw, j, X.n,
                                    trk ? O.sn + si : nullptr, O.lastcol ? O.last + si : nullptr);
                }
            }
        }
    }
    if (SPECIAL) {
        STITCH_UNROLL
        for (int k = 0; k < STRIP; ++k) {
            if (k >= nv) Sn[k] = Sup[k];                       // row m is finished per contig; padding keeps its value
            if (k > nv || (k == nv && !has_m)) st.D6[k] = Dup[k];
        }
        if (has_m) {
            PkRowM rm; rm.diag = st.A[nv]; rm.D6 = st.D6[nv]; rm.jp = st.jp[nv]; rm.I = I_m; rm.fl = TB ? st.fl[nv] : 0u; rm.iext = iext_m;
            X.team.peer(S.stash, a % X.team.size)[a] = rm;   // to the CTA that finishes contig a
            if (QUIET) S.DmKey[a] = rm.D6;
        }
    }
    *reinterpret_cast<int4 *>(X.Sst + tile * ST + lane * 4) = make_int4(Sn[0], Sn[1], Sn[2], Sn[3]);
    *reinterpret_cast<int4 *>(X.Sst + tile * ST + 128 + lane * 4) = make_int4(Sn[4], Sn[5], Sn[6], Sn[7]);
    *reinterpret_cast<int4 *>(X.Dst + tile * ST + lane * 4) = make_int4(st.D6[0], st.D6[1], st.D6[2], st.D6[3]);
    *reinterpret_cast<int4 *>(X.Dst + tile * ST + 128 + lane * 4) = make_int4(st.D6[4], st.D6[5], st.D6[6], st.D6[7]);
    colmax = __reduce_max_sync(FULL, colmax);   // (REDUX.MAX: one instruction instead of a five-step shuffle butterfly on the tile's critical path)
    if (lane < X.team.size) X.team.peer(S.tilemax, lane)[tile] = colmax;   // every CTA of the team holds the whole tile table
    if (QUIET) {   // is the tile in the closed form of column j?  (S first: the cheap test, and the one that usually fails)
        const PkQuiet qn = *qz.Qn;
        bool ok = true;
        if (!SPECIAL && !TB) ok = (notq & pk.NPM) == 0;   // jp & NPM is the closed form's key of the cell (pk_quiet_next)
        else {
            STITCH_UNROLL
            for (int k = 0; k < STRIP; ++k)
                if (!SPECIAL || k < nv) ok = ok && Sn[k] == (x[k] == pc.q ? qn.bk[0] : qn.bk[1]);
        }
        if (!__all_sync(FULL, ok)) return false;
        STITCH_UNROLL
        for (int k = 0; k < STRIP; ++k)
            if (!SPECIAL || k < nv) ok = ok && (st.D6[k] == pk_quiet_D_dev(pk, qn, x[k], qz.yq) || (st.D6[k] >> pk.SH) <= qz.deadrel);
        return __all_sync(FULL, ok);
    }
    return false;
}

// Halos of the current state into parity slot `slot` (before the first column computed from it).
template <int W>
__device__ void pk_init_halos(const PackCtx &X, PackSmem &S, uint32_t slot) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t GW = X.team.size * W, gw = X.team.rank * W + warp;
    const uint32_t Weff = X.NT < GW ? X.NT : GW;
    if (gw >= 1 && gw < Weff && lane < 9) {
        const uint32_t t_lo = X.win_lo + (uint32_t)((uint64_t)X.NT * gw / Weff);
        const uint32_t hl = lane == 0 ? 30u : 31u, hk = lane == 0 ? (uint32_t)STRIP - 1 : lane - 1;
        const int32_t *tp = pk_tile_ptr(X, t_lo - 1) + (hk >> 2) * 128u + hl * 4u + (hk & 3u);
        S.haloS[(slot * W + warp) * 9 + lane] = tp[0];
        if (lane >= 1) S.haloD[(slot * W + warp) * 8 + lane - 1] = tp[TILE];
    }
}

// Phases T and F of one column.  On entry S.Jc (and S.Jw for the traceback variant), S.Sm/slm/tbm/SmKey of
// column j-1 and the halos of parity (j-1)&1 are set and the CTA is synchronised; on exit S.cm/cml/cmk,
// S.Sm/slm/tbm/SmKey describe column j and the CTA is synchronised.
__device__ __forceinline__ uint32_t pk_base_bit(uint8_t b) { return b == 'A' ? 1u : b == 'C' ? 2u : b == 'G' ? 4u : b == 'T' ? 8u : 16u; }

template <int W, bool TB, bool QUIET = false>
__device__ void pk_column(const PackCtx &X, PackSmem &S, const PCol &pc, int32_t r0pkey, int32_t cr1key, uint32_t j,
                          const PkColOut &O, const uint8_t *yq = nullptr, PkColStat *cs = nullptr) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = __shfl_sync(FULL, tid >> 5, 0);   // (a shuffle from lane 0: the compiler then treats the warp index and what derives from it as warp-uniform)
    const PK &pk = X.pk;
    const Scoring &sc = X.sc;
    const uint32_t NT = X.NT, C = X.C, par = j & 1u;
    const Team &team = X.team;
    const uint32_t GW = team.size * W, gw = team.rank * W + warp;
    const uint32_t Weff = NT < GW ? NT : GW;
    const PkQuiet *Qp = S.Q + (par ^ 1u) * S.cmax, *Qn = S.Q + par * S.cmax;
    const bool cst = cs && cs->timing;   // per-phase cycle counters (STITCH_DEBUG_STATS only: 64-bit shared-memory atomics per warp and column)
    const long long c0 = cst ? clock64() : 0;
    if (gw < Weff && X.win_lo + (uint32_t)((uint64_t)NT * (gw + 1) / Weff) > X.skip_below) {
        const uint32_t t_lo = X.win_lo + (uint32_t)((uint64_t)NT * gw / Weff), t_hi = X.win_lo + (uint32_t)((uint64_t)NT * (gw + 1) / Weff);
        const int32_t *hS = S.haloS + ((par ^ 1u) * W + warp) * 9, *hD = S.haloD + ((par ^ 1u) * W + warp) * 8;
        int32_t prev_exit = 0, prev_s7 = 0; uint32_t prev_exit_open = 0;
        // bulk copies bring the next tiles (S keys, D keys, bases) into this warp's shared-memory ring while a tile is computed:
        // one elected lane issues a tile (cp.async.bulk: the state and the bases are contiguous), every lane waits on the slot's
        // mbarrier
        const bool staged = X.staged;
        PkQuietArgs qz; qz.mat = false; qz.prev_skipped = false; qz.Qp = Qp; qz.Qn = Qn; qz.yq = yq; qz.deadrel = QUIET ? pk_deadrel(sc) : 0;
        // quiet tiles: the plan of this warp's chunk for this column, from the flags of column j-1 (a warp owns the flags
        // of its chunk), one tile per lane: 0 = skip, 1 = compute from the materialised closed form, 2 = load + compute.
        // The first and last tile of a chunk and of a contig are always computed (halos, row 1, row m).  A skipped tile
        // only contributes its best key (the better of the closed form's two keys among the bases it holds).  The tiles that
        // are computed go, in order, into this warp's slice of S.clist as descriptors: bits 0..12 chunk-relative tile, bit 13
        // the tile may turn quiet, bits 14..15 the mode.
        const bool edge = QUIET && X.quiet_edge;
        const uint32_t lq0 = (edge && gw > 0) ? S.haloF[(par ^ 1u) * W + warp] : 0u;   // last tile of the previous chunk was quiet at column j-1
        uint16_t *cl = S.clist + t_lo;
        uint32_t ncomp = t_hi - t_lo, nskipped = 0, mode_first = 2u, mode_last = 2u;
        if (QUIET) {
            const uint32_t mb = pk_base_bit(pc.q);
            ncomp = 0;
            for (uint32_t base = t_lo; base < t_hi; base += 32) {
                const uint32_t t = base + lane;
                uint32_t mode = 2, tbv = 0;
                if (t < t_hi) {
                    tbv = S.tb[t];
                    if (tbv & PackSmem::TB_Q) {
                        const uint32_t a_t = X.owner[t];
                        const uint32_t tic = t - X.ent[a_t].tile_start;
                        const PkQuiet &qn = Qn[a_t];
                        // (the first tile of a contig has no upper neighbour but four more conditions: stay_first; the first tile
                        // of a chunk reads the flag its neighbour chunk's owner published with the halo)
                        const bool left_ok = t != t_lo ? (S.tb[t - 1] & PackSmem::TB_Q) != 0 : lq0 != 0u;
                        const bool lastt = tic + 1 == X.ent[a_t].ntiles;
                        mode = ((!lastt || (X.quiet_last && tic != 0)) && (edge || (t != t_lo && t + 1 != t_hi)) &&
                                (tic == 0 ? qn.stay_first != 0 : (left_ok && qn.stay))) ? 0u : 1u;
                        if (mode == 0) {
                            const uint32_t tm = tbv & PackSmem::TB_MASK;
                            S.tilemax[t] = tm == 0u ? pk.NEGKEY : ((tm & mb) ? ((tm & ~mb) ? pk_max(qn.bk[0], qn.bk[1]) : qn.bk[0]) : qn.bk[1]);
                            if (lastt) {   // row m: the candidates the tile would have stashed for the per-contig finish
                                const ContigEntry &e = X.ent[a_t];
                                int32_t dmn;
                                S.stash[a_t] = pk_quiet_rowm(pk, pc, Qp[a_t], S.Jc[a_t], S.SmKey[a_t], S.DmKey[a_t],
                                                             X.bases[e.seq_off + e.m - 2] == yq[1] ? 0 : 1, X.bases[e.seq_off + e.m - 1] == pc.q, &dmn);
                                S.DmKey[a_t] = dmn;
                            }
                        }
                    }
                }
                const bool comp = t < t_hi && mode != 0u;
                const uint32_t m = __ballot_sync(FULL, comp);
                if (comp) cl[ncomp + __popc(m & ((1u << lane) - 1u))] = (uint16_t)((t - t_lo) | ((tbv & PackSmem::TB_CANQ) ? 0x2000u : 0u) | (mode << 14));
                ncomp += __popc(m);
                if (base == t_lo) mode_first = __shfl_sync(FULL, mode, 0);
                if (t_hi - 1u - base < 32u) mode_last = __shfl_sync(FULL, mode, (int)(t_hi - 1u - base));
            }
            nskipped = (t_hi - t_lo) - ncomp;
            __syncwarp();
            if (lq0 && mode_first != 0u) {
                // the chunk's first tile is computed and its upper neighbour (another warp's tile) is quiet, its memory possibly
                // stale: the 17-key halo of column j-1 comes from the closed form instead of the published copy
                const uint32_t a0 = X.owner[t_lo];
                const uint32_t tic0 = t_lo - X.ent[a0].tile_start;
                if (tic0 != 0 && lane < 9) {
                    const uint8_t xb = X.bases[X.ent[a0].seq_off + tic0 * TILE - STRIP + lane - 1];   // rows hrow0-1 .. hrow0+7
                    const PkQuiet qp = Qp[a0];
                    S.haloS[((par ^ 1u) * W + warp) * 9 + lane] = xb == yq[1] ? qp.bk[0] : qp.bk[1];
                    if (lane >= 1) S.haloD[((par ^ 1u) * W + warp) * 8 + lane - 1] = pk_quiet_D_dev(pk, qp, xb, yq + 1);
                }
                __syncwarp();
            }
        }
        auto desc_at = [&](uint32_t k) -> uint32_t { return QUIET ? (uint32_t)cl[k] : (k | (2u << 14)); };
        // contig of the current tile (reloaded only when the chunk crosses into the next contig)
        uint32_t a = X.owner ? X.owner[t_lo] : 0u;
        ContigEntry en = X.ent[a];
        int32_t Jc = S.Jc[a];
        // software pipeline: a ring of `depth` stage slots per warp; while tile k is computed the next depth - 1 computed tiles are
        // in flight, so that a tile's HBM / L2 latency is covered by several tiles of arithmetic
        const uint32_t dmask = X.stage_depth - 1u;
        unsigned char *stgw = S.stage + (size_t)warp * X.stage_depth * PackSmem::STAGE_BYTES;
        const uint32_t stg_s = smem_u32(stgw), bar_s = smem_u32(S.mbar + warp * 4u);
        uint32_t phase = staged ? S.mphase[warp] : 0u;
        uint32_t issued = 0;
        auto issue = [&](uint32_t k) {
            if (lane == 0) {
                const uint32_t d = desc_at(k), t = t_lo + (d & 0x1fffu), slot = k & dmask;
                const uint32_t dst = stg_s + slot * PackSmem::STAGE_BYTES, bar = bar_s + slot * 8u;
                const bool full = (d >> 14) == 2u;   // (a materialised tile only needs its bases)
                mbar_expect_tx(bar, full ? PackSmem::STAGE_BYTES : PackSmem::STAGE_PRE + (uint32_t)TILE);
                if (full) bulk_g2s(dst, X.Sst + (size_t)t * ST, 2 * TILE * 4, bar);
                bulk_g2s(dst + 2 * TILE * 4, X.tbases + (size_t)t * TILE - PackSmem::STAGE_PRE, PackSmem::STAGE_PRE + (uint32_t)TILE, bar);
            }
        };
        if (staged) for (; issued < dmask && issued < ncomp; ++issued) issue(issued);
        uint32_t prev_tile = 0xfffffffeu;
        for (uint32_t k = 0; k < ncomp; ++k) {
            const uint32_t d = desc_at(k);
            const uint32_t tile = t_lo + (d & 0x1fffu);
            const bool is_load = (d >> 14) == 2u;
            qz.prev_skipped = QUIET && tile != t_lo && prev_tile + 1u != tile;
            prev_tile = tile;
            if (tile >= en.tile_start + en.ntiles) { a = X.owner[tile]; en = X.ent[a]; Jc = S.Jc[a]; }
            const unsigned char *stg = nullptr;
            if (staged) {
                // (the slot tile k + depth - 1 goes into was read by tile k - 1: every lane is past those reads, the tile ends in
                // warp-wide votes)
                if (issued < ncomp) { issue(issued); ++issued; }
                const uint32_t slot = k & dmask;
                mbar_wait(bar_s + slot * 8u, (phase >> slot) & 1u);
                phase ^= 1u << slot;
                stg = stgw + slot * PackSmem::STAGE_BYTES;
            }
            qz.mat = QUIET && !is_load;
            qz.Qp = Qp + a; qz.Qn = Qn + a;
            const uint32_t tic = tile - en.tile_start;
            bool qnow;
            if (tic == 0 || tic + 1 == en.ntiles)
                qnow = pk_tile<true, TB, QUIET>(X, pc, S, O, j, tile, lane, r0pkey, cr1key, tile == t_lo, hS, hD, prev_exit, prev_exit_open, prev_s7, stg, a, en, Jc, qz);
            else
                qnow = pk_tile<false, TB, QUIET>(X, pc, S, O, j, tile, lane, r0pkey, cr1key, tile == t_lo, hS, hD, prev_exit, prev_exit_open, prev_s7, stg, a, en, Jc, qz);
            if (QUIET && lane == 0) {
                const uint32_t tbv = S.tb[tile];
                S.tb[tile] = (uint8_t)((tbv & ~PackSmem::TB_Q) | ((qnow && (d & 0x2000u)) ? PackSmem::TB_Q : 0u));
            }
        }
        if (staged && lane == 0) S.mphase[warp] = phase;
        if (cs && lane == 0) { if (nskipped) atomicAdd(&cs->skipped, nskipped); if (cst) { const unsigned long long dt = (unsigned long long)(clock64() - c0); atomicAdd(&cs->t_busy, dt); cs->t_warp[warp] += dt; cs->n_warp[warp] += ncomp; } }
        if (QUIET) __syncwarp();   // (lane 0's flag of the last tile)
        if (QUIET && gw + 1 < Weff && lane == 0) S.haloF[par * W + warp + 1] = (S.tb[t_hi - 1] & PackSmem::TB_Q) ? 1u : 0u;
        if (gw + 1 < Weff && (!QUIET || mode_last != 0u)) {
            // publish the halo of the next chunk for the next column (the next CTA's warp 0 after our last warp); a skipped
            // last tile publishes only its flag (above): its memory is stale and the consumer uses the closed form
            const bool local = warp + 1 < (uint32_t)W;
            const uint32_t slot = local ? warp + 1 : 0u;
            int32_t *nS = (local ? S.haloS : team.peer(S.haloS, team.rank + 1)) + (par * W + slot) * 9;
            int32_t *nD = (local ? S.haloD : team.peer(S.haloD, team.rank + 1)) + (par * W + slot) * 8;
            // (read back from the state this warp has just written: keeps 16 registers out of the tile loop)
            const int32_t *tp = X.Sst + (t_hi - 1) * ST;
            if (lane == 31) {
                const int4 s0 = *reinterpret_cast<const int4 *>(tp + 31 * 4), s1 = *reinterpret_cast<const int4 *>(tp + 128 + 31 * 4);
                const int4 d0 = *reinterpret_cast<const int4 *>(tp + TILE + 31 * 4), d1 = *reinterpret_cast<const int4 *>(tp + TILE + 128 + 31 * 4);
                nS[1] = s0.x; nS[2] = s0.y; nS[3] = s0.z; nS[4] = s0.w; nS[5] = s1.x; nS[6] = s1.y; nS[7] = s1.z; nS[8] = s1.w;
                nD[0] = d0.x; nD[1] = d0.y; nD[2] = d0.z; nD[3] = d0.w; nD[4] = d1.x; nD[5] = d1.y; nD[6] = d1.z; nD[7] = d1.w;
            }
            if (lane == 30) nS[0] = tp[128 + 30 * 4 + 3];
        }
    }
    if (X.staged) fence_async_proxy();   // this column's state stores before the next column's bulk copies of the same tiles
    team.sync();
    if (X.no_finish) return;   // cone re-fill: row m is outside the window, nothing reads the column best

    const long long c1 = cst ? clock64() : 0;
    // ---- per contig: tracker + row m + column best (contig a on CTA a % size) ----
    const Row0 r0 = row0_at(sc, j, X.n);
    // Four contigs per warp at a time, one per group of 8 lanes (the groups' memory latencies overlap and the four
    // serial row-m finishes run side by side on the groups' first lanes).
    const uint32_t grp = lane >> 3, gl = lane & 7u;
    for (uint32_t slot = warp * 4u + grp; __any_sync(FULL, team.rank + team.size * slot < C); slot += (uint32_t)W * 4u) {
        const bool act = team.rank + team.size * slot < C;
        const uint32_t a = act ? team.rank + team.size * slot : 0u;   // idle groups shadow contig 0 (no writes)
        const ContigEntry en = X.ent[a];
        int32_t kmax = pk.NEGKEY;
        for (uint32_t t = gl; t < en.ntiles; t += 8) kmax = pk_max(kmax, S.tilemax[en.tile_start + t]);
        STITCH_UNROLL
        for (int d = 4; d >= 1; d >>= 1) kmax = pk_max(kmax, __shfl_xor_sync(FULL, kmax, d));
        const int32_t smax = pk_rel(pk, kmax);
        if (cst && tid == 0) cs->t_fa += (unsigned long long)(clock64() - c1);
        // first row (< m) whose S has the best score (column best, SCA:680-687) / equals the best key (tracker, SCA:411-416)
        uint32_t frow = 0, trow = 0; int32_t fkey = 0;
        STITCH_UNROLL
        for (int pass = 0; pass < (TB ? 2 : 1); ++pass) {
            const bool full = pass == 1;
            uint32_t ft = 0xffffffffu;
            for (uint32_t t = gl; t < en.ntiles; t += 8) {
                const int32_t tm = S.tilemax[en.tile_start + t];
                if (full ? tm == kmax : pk_rel(pk, tm) == smax) { ft = t; break; }
            }
            STITCH_UNROLL
            for (int d = 4; d >= 1; d >>= 1) { const uint32_t o = __shfl_xor_sync(FULL, ft, d); ft = o < ft ? o : ft; }
            uint32_t row = 0xffffffffu; int32_t key = 0;
            if (en.m >= 2 && ft != 0xffffffffu) {
                const uint32_t tile = en.tile_start + ft;
                const bool quiet_tile = QUIET && (S.tb[tile] & PackSmem::TB_Q);   // (its memory may be stale): the closed form of column j
                const int32_t *tp = quiet_tile ? nullptr : pk_tile_ptr(X, tile);
                // the group's lane g scans strips 4g .. 4g+3 of the tile (32 consecutive rows), last strip first so that the
                // first matching row is what remains
                STITCH_UNROLL
                for (int u = 3; u >= 0; --u) {
                    const uint32_t sl_ = gl * 4u + (uint32_t)u;
                    int32_t sk[STRIP];
                    if (quiet_tile) {
                        uint8_t xq[STRIP];
                        unpack8(*reinterpret_cast<const uint2 *>(X.bases + en.seq_off + ft * TILE + sl_ * STRIP), xq);
                        const PkQuiet &qn = Qn[a];
                        STITCH_UNROLL
                        for (int k = 0; k < STRIP; ++k) sk[k] = xq[k] == pc.q ? qn.bk[0] : qn.bk[1];
                    } else {
                        const int4 s0 = pk_ld_state(X, tp + sl_ * 4);
                        const int4 s1 = pk_ld_state(X, tp + 128 + sl_ * 4);
                        sk[0] = s0.x; sk[1] = s0.y; sk[2] = s0.z; sk[3] = s0.w; sk[4] = s1.x; sk[5] = s1.y; sk[6] = s1.z; sk[7] = s1.w;
                    }
                    STITCH_UNROLL
                    for (int k = STRIP - 1; k >= 0; --k) {
                        const uint32_t i = ft * TILE + sl_ * STRIP + (uint32_t)k + 1;
                        if (i < en.m && (full ? sk[k] == kmax : pk_rel(pk, sk[k]) == smax)) { row = i; key = sk[k]; }
                    }
                }
            }
            uint32_t best = row;
            STITCH_UNROLL
            for (int d = 4; d >= 1; d >>= 1) { const uint32_t o = __shfl_xor_sync(FULL, best, d); best = o < best ? o : best; }
            const uint32_t hit = (__ballot_sync(FULL, row == best) >> (grp * 8u)) & 0xffu;   // never empty: the best lane itself
            key = __shfl_sync(FULL, key, grp * 8u + (uint32_t)__ffs(hit) - 1u);
            if (en.m >= 2) { if (full) trow = best; else { frow = best; fkey = key; } }
        }
        if (cst && tid == 0) cs->t_f1 += (unsigned long long)(clock64() - c1);
        if (gl == 0 && act) {
            CmPart rows; cm_init(rows);
            XsPart tr; xs_init(tr);
            if (en.m >= 2) {
                rows.S = pc.B + smax; rows.row = frow; rows.sl = pk_len(pk, fkey); rows.valid = 1;
                if (sc.xs != MIN_SCORE) { tr.t = pc.B + smax + sc.xs; tr.len = pk_len(pk, kmax); tr.row = TB ? trow : 1u; }
            }
            const PkRowM stash = S.stash[a];
            JumpInfo Jw; Jw.score = 0; Jw.len = 0; Jw.idx = 0; Jw.from = 0;
            if (TB) Jw = S.Jw[a];
            const PkRowMOut fo = pk_finish_rowm(pk, pc, sc, stash, tr, r0, Jw, X.bases[en.seq_off + en.m - 1] == pc.q, en.contig_idx, en.m, j);
            const RowMOut &ro = fo.ro;
            const uint32_t r = en.m - 1;
            const uint32_t mt = en.tile_start + r / TILE, ml = (r % TILE) / STRIP, mk = r % STRIP;
            pk_tile_ptr(X, mt)[(mk >> 2) * 128u + ml * 4u + (mk & 3u)] = fo.skey;
            if (TB) {
                if (O.tb_col) O.tb_col[mt * TILE + ml * STRIP + mk] = (uint8_t)fo.tbbyte;
                if (O.colrec_col) {
                    ColRec cr; cr.jscore = Jw.score; cr.jlen = Jw.len; cr.jidx = Jw.idx; cr.jfrom = Jw.from; cr.lx = ro.lx;
                    cr.pad0 = cr.pad1 = cr.pad2 = 0;
                    O.colrec_col[a] = cr;
                }
                const uint32_t p = state_index(mt, ml, mk);
                if (O.track && ro.c.S >= O.track_thr) sn_update(sc, O.sn[p], ro.c.S, ro.c.sl, ro.c.idx, j, X.n);
                if (O.lastcol) {
                    LastCell lc; lc.S = ro.c.S; lc.I = pk_abs(pk, pc.B, stash.I); lc.sl = ro.c.sl; lc.il = pk_len(pk, stash.I);
                    lc.idx = ro.c.idx; lc.from = ro.c.from; lc.s_tb = (uint8_t)ro.s_tb; lc.i_tb = 0;
                    lc.flags = (uint8_t)((stash.iext ? 1 : 0) | ((stash.fl & 1u) ? 2 : 0)); lc.pad = 0; lc.pad2 = 0;
                    O.last[p] = lc;
                }
            }
            CmPart cmv; cm_init(cmv);
            cm_add(cmv, r0.S, r0.sl, 0);
            cmv = cm_merge(cmv, rows);
            CmPart top; top.S = ro.c.S; top.row = en.m; top.sl = ro.c.sl; top.valid = 1;
            cmv = cm_merge(cmv, top);
            for (uint32_t r = 0; r < team.size; ++r) {   // every CTA of the team keeps the per-contig tables
                team.peer(S.cm, r)[a] = cmv.S; team.peer(S.cmk, r)[a] = cmv.row; team.peer(S.cml, r)[a] = cmv.sl;
                team.peer(S.Sm, r)[a] = ro.c.S; team.peer(S.slm, r)[a] = ro.c.sl; team.peer(S.tbm, r)[a] = ro.s_tb;
                team.peer(S.SmKey, r)[a] = fo.skey;
            }
        }
    }
    if (cst && tid == 0) cs->t_f2 += (unsigned long long)(clock64() - c1);
    if (X.staged) fence_async_proxy();   // (row m's key)
    team.sync();
    if (cst && tid == 0) { cs->t_tiles += (unsigned long long)(c1 - c0); cs->t_finish += (unsigned long long)(clock64() - c1); }
}

// Column 0 (SCA:97-186) of the contigs of X into the packed state (base B_0 = 0) + row-m summaries.
template <int W>
__device__ void pk_state_init0(const PackCtx &X, PackSmem &S) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr uint32_t T = W * 32;
    const PK &pk = X.pk;
    for (uint32_t p = X.own_lo * ST + tid; p < X.own_hi * ST; p += T) X.Sst[p] = (p % ST) < (uint32_t)TILE ? pk.NEGKEY : pk.NEGKEY + pk.PD6;
    __syncthreads();
    for (uint32_t tile = X.own_lo + warp; tile < X.own_hi; tile += W) {
        const ContigEntry en = X.ent[X.owner ? X.owner[tile] : 0u];
        const uint32_t tic = tile - en.tile_start;
        STITCH_UNROLL
        for (int k = 0; k < STRIP; ++k) {
            const uint32_t i = tic * TILE + lane * STRIP + (uint32_t)k + 1;
            if (i <= en.m) {
                const Col0 c0 = col0_at(X.sc, i, en.m);
                X.Sst[pk_sidx(tile, lane, (uint32_t)k)] = pk_from_wide(pk, 0, c0.S, c0.sl, 0);
            }
        }
    }
    for (uint32_t a = tid; a < X.C; a += T) {
        const ContigEntry en = X.ent[a];
        S.cm[a] = 0; S.cml[a] = 0; S.cmk[a] = 0;
        const Col0 cmm = col0_at(X.sc, en.m, en.m);
        S.Sm[a] = cmm.S; S.slm[a] = cmm.sl; S.tbm[a] = cmm.s_tb;
        S.SmKey[a] = pk_from_wide(pk, 0, cmm.S, cmm.sl, 0);
        S.DmKey[a] = pk.NEGKEY + pk.PD6;
    }
    if (X.staged) fence_async_proxy();
    X.team.sync();
}

// The packed state of column j0 from a checkpoint: a raw copy of the state arrays (keys relative to the base of
// column j0, which the caller passes for the row-m key); `ck` points at the first tile of X.
template <int W>
__device__ void pk_state_from_ck(const PackCtx &X, PackSmem &S, const int32_t *ck, const CkSum *sums, int32_t Bj0) {
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t T = W * 32;
    const PK &pk = X.pk;
    const int4 *src = reinterpret_cast<const int4 *>(ck);
    int4 *dst = reinterpret_cast<int4 *>(X.Sst);
    for (uint32_t idx = X.own_lo * (ST / 4) + tid; idx < X.own_hi * (ST / 4); idx += T) dst[idx] = __ldcs(src + idx);
    for (uint32_t a = tid; a < X.C; a += T) {
        const CkSum cs = sums[a];
        S.cm[a] = 0; S.cml[a] = 0; S.cmk[a] = 0;
        S.Sm[a] = cs.Sm; S.slm[a] = cs.slm; S.tbm[a] = cs.tbm;
        S.SmKey[a] = pk_from_wide(pk, Bj0, cs.Sm, cs.slm, 0);
    }
    if (X.staged) fence_async_proxy();
    X.team.sync();
}

// Checkpoint of the current packed state: a raw copy (8 bytes per cell).
template <int W>
__device__ void pk_write_ck(const PackCtx &X, PackSmem &S, int32_t *dck, CkSum *dsum) {
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t T = W * 32;
    const int4 *src = reinterpret_cast<const int4 *>(X.Sst);
    int4 *dst = reinterpret_cast<int4 *>(dck);
    // streaming store: checkpoints are read back once, much later; keep the L2 for the rolling state
    for (uint32_t idx = X.own_lo * (ST / 4) + tid; idx < X.own_hi * (ST / 4); idx += T)
        __stcs(dst + idx, pk_ld_state(X, reinterpret_cast<const int32_t *>(src + idx)));
    if (X.team.rank == 0)
        for (uint32_t a = tid; a < X.C; a += T) {
            CkSum cs; cs.Sm = S.Sm[a]; cs.slm = S.slm[a]; cs.tbm = S.tbm[a]; cs.pad = 0;
            dsum[a] = cs;
        }
}

// ---------------------------------------------------------------------------------------------
// tail: columns (j0, n] again in the traceback variant
struct PkColConst { PCol pc; int32_t r0pkey, cr1key; uint8_t yq[PKQ_L + 2]; };   // yq[k] = y_{j-k} (bulk pass), 0 = none
__device__ __forceinline__ PkColConst pk_col_const(const PK &pk, const Scoring &sc, int32_t B, int32_t Bprev, uint32_t j, uint32_t n, uint8_t q) {
    PkColConst c;
    c.pc = pk_col(pk, sc, B, Bprev, j, n, q);
    const Row0 r0 = row0_at(sc, j, n), r0p = row0_at(sc, j - 1, n);
    c.r0pkey = pk_from_wide(pk, Bprev, r0p.S, r0p.sl, 0);
    c.cr1key = pk_carry_row1(pk, c.pc, sc, r0);
    return c;
}

// ---------------------------------------------------------------------------------------------
// Closed form of a quiet tile of contig `a` in column j (dp_packed.h) and whether quiet tiles of that contig stay quiet;
// S.Jc / S.SmKey / S.tbm of the contig are up to date.  `allow` = the column may skip tiles at all.
__device__ __forceinline__ void pk_quiet_contig(const PackCtx &X, PackSmem &S, uint32_t a, uint32_t j, const PkColConst &ccl, bool allow) {
    const uint32_t par = j & 1u;
    PkFirstIn fi;   // what row 1 of this contig sees beyond an ordinary row
    fi.r0pkey = ccl.r0pkey; fi.cr1key = ccl.cr1key; fi.wbase = pk_wbase(X.pk, S.SmKey[a]);
    fi.wrap = X.ent[a].circular && S.tbm[a] != TB_XCLIP_SUFFIX;
    fi.yc1 = X.yclip_mode ? pk_key(X.pk, (int64_t)X.sc.yp + X.sc.o + (int64_t)X.sc.e - ccl.pc.B, PP_YC, col0_slen(X.sc, 1, X.ent[a].m)) : X.pk.NEGKEY;
    S.Q[par * S.cmax + a] = pk_quiet_next(X.pk, X.sc, ccl.pc, S.Jc[a], S.Q[(par ^ 1u) * S.cmax + a], allow && pk_base_bit(ccl.pc.q) != 16u,
                                          X.quiet_first ? &fi : nullptr);
}

// Replays the jump records and column bases the bulk pass left (colrec, gcol).
template <int W>
__device__ __forceinline__ void pk_replay_consts(const PackCtx &X, PackSmem &S, const ColRec *colrec, const int32_t *gcol,
                                                 const uint8_t *read, uint32_t j, uint32_t a_global, uint32_t Cglobal, PkColConst *s_cc,
                                                 bool tail_quiet = false, bool track = false, int32_t track_thr = 0) {
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t T = W * 32;
    const int32_t B = gcol[j - 1], Bprev = j >= 2 ? gcol[j - 2] : 0;
    PCol pcl; pcl.B = B; pcl.delta = B - Bprev;
    for (uint32_t a = tid; a < X.C; a += T) {
        const ColRec cr = colrec[(uint64_t)j * Cglobal + a_global + a];
        JumpInfo J; J.score = cr.jscore; J.len = cr.jlen; J.idx = cr.jidx; J.from = cr.jfrom;
        S.Jw[a] = J;
        S.Jc[a] = pk_jc(X.pk, pcl, J.score, J.len);
        if (tail_quiet) {
            // quiet tiles in the tail: a skipped tile must not hold a cell that can still set a final y-suffix tracker
            // (SCA:432-447), and column n leaves the records of every cell for the end-of-read fix-up
            pk_quiet_contig(X, S, a, j, pk_col_const(X.pk, X.sc, B, Bprev, j, X.n, read[j - 1]), j != X.n);
            PkQuiet &qn = S.Q[(j & 1u) * S.cmax + a];
            if (track && pk_abs(X.pk, B, pk_max(qn.bk[0], qn.bk[1])) >= track_thr) { qn.stay = 0; qn.stay_first = 0; }
        }
    }
    if (tid == 0) {
        *s_cc = pk_col_const(X.pk, X.sc, B, Bprev, j, X.n, read[j - 1]);
        STITCH_UNROLL
        for (int k = 0; k < PKQ_L + 2; ++k) s_cc->yq[k] = j >= 1u + (uint32_t)k ? read[j - 1 - (uint32_t)k] : (uint8_t)0;
    }
}

// Late checkpoints.  The tail restarts from the last checkpoint before the first column that can hold a final y-suffix
// tracker, typically a few tens of columns before the end of the read; with regular checkpoints alone it re-runs up to K more
// columns in the (slower) traceback variant.  PK_LATE_CKS extra checkpoints at columns n - PK_LATE_STEP * k (k = 1 ..
// PK_LATE_CKS, unless that is a regular checkpoint column) are stored after the regular ones.
constexpr uint32_t PK_LATE_CKS = 3, PK_LATE_STEP = 32;
__host__ __device__ inline uint32_t pk_regular_cks(uint32_t n, uint32_t K) { return (n - 1) / K; }   // columns K, 2K, ... < n
// k (1 .. PK_LATE_CKS) when column j is a late checkpoint column of a read of n columns, else 0
__device__ __forceinline__ uint32_t pk_late_ck(uint32_t j, uint32_t n, uint32_t K) {
    if (j >= n || j % K == 0u) return 0u;
    const uint32_t d = n - j;
    return (d % PK_LATE_STEP == 0u && d / PK_LATE_STEP <= PK_LATE_CKS) ? d / PK_LATE_STEP : 0u;
}
// index (among the read's checkpoints) of the one at column j0 > 0
__device__ __forceinline__ uint32_t pk_ck_index(uint32_t j0, uint32_t n, uint32_t K) {
    return j0 % K == 0u ? j0 / K - 1u : pk_regular_cks(n, K) + (n - j0) / PK_LATE_STEP - 1u;
}

// The tail of one job on the CTA that ran its bulk pass (s_cc: one shared PkColConst slot).
template <int W>
__device__ void pk_tail(const Params &P, const JobDesc &jd, const LayoutDesc &ld, PackCtx &X, PackSmem &S, PkColConst *s_cc, uint32_t j0,
                        int32_t track_thr) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const Scoring &sc = X.sc;
    const uint32_t C = ld.C, PM = ld.PM, n = jd.n, K = P.K;
    ColRec *colrec = P.colrec + jd.colrec_off;
    const int32_t *gcol = P.gcol + jd.gcol_off;
    const uint8_t *read = P.reads + jd.read_off;
    SnRec *sn = P.sn + jd.cell_off;
    const bool tracked = P.tracked_mode != 0;
    if (j0 == 0) pk_state_init0<W>(X, S);
    else pk_state_from_ck<W>(X, S, P.pck + jd.ck_off + (uint64_t)pk_ck_index(j0, n, K) * 2 * PM,
                             P.ck_sum + jd.cksum_off + (uint64_t)pk_ck_index(j0, n, K) * C, gcol[j0 - 1]);
    if (tracked) {   // trackers start from column 0 (SCA:179-183)
        for (uint32_t tile = X.team.rank * W + warp; tile < X.NT; tile += X.team.size * W) {
            const ContigEntry en = X.ent[X.owner[tile]];
            const uint32_t tic = tile - en.tile_start;
            STITCH_UNROLL
            for (int k = 0; k < STRIP; ++k) {
                const uint32_t i = tic * TILE + lane * STRIP + (uint32_t)k + 1;
                if (i <= en.m) {
                    const Col0 c0 = col0_at(sc, i, en.m);
                    sn[state_index(tile, lane, (uint32_t)k)] = sn_init(sc, c0.S, c0.sl, en.contig_idx, n);
                }
            }
        }
    }
    pk_init_halos<W>(X, S, j0 & 1u);
    // quiet tiles in the tail (same closed form as in the bulk pass; no tile is quiet at the checkpoint): the row-m D keys come
    // from the state just restored
    const bool tail_quiet = X.quiet && P.quiet_tail != 0 && X.team.size == 1;
    if (tail_quiet) {
        for (uint32_t a = tid; a < C; a += W * 32) {
            const ContigEntry &e = X.ent[a];
            const uint32_t r = e.m - 1, mt = e.tile_start + r / TILE, ml = (r % TILE) / STRIP, mk = r % STRIP;
            S.DmKey[a] = X.Dst[mt * ST + (mk >> 2) * 128u + ml * 4u + (mk & 3u)];
            S.Q[a] = pk_quiet_init(X.pk); S.Q[S.cmax + a] = pk_quiet_init(X.pk);
        }
        for (uint32_t t = tid; t < X.NT; t += W * 32) S.tb[t] = (uint8_t)(S.tb[t] & (PackSmem::TB_MASK | PackSmem::TB_CANQ));
        if (tid < 2 * W) S.haloF[tid] = 0;
    }
    X.team.sync();   // trackers of every row initialised before any CTA updates them
    for (uint32_t j = j0 + 1; j <= n; ++j) {
        pk_replay_consts<W>(X, S, colrec, gcol, read, j, 0, C, s_cc, tail_quiet, tracked, track_thr);
        __syncthreads();
        const PkColConst cc = *s_cc;
        PkColOut O; O.tb_col = nullptr; O.colrec_col = colrec + (uint64_t)j * C; O.sn = sn; O.last = P.last + jd.cell_off;
        O.track = tracked; O.lastcol = j == n; O.track_thr = track_thr;
        if (tail_quiet) pk_column<W, true, true>(X, S, cc.pc, cc.r0pkey, cc.cr1key, j, O, cc.yq);
        else pk_column<W, true>(X, S, cc.pc, cc.r0pkey, cc.cr1key, j, O);
    }
}

// Base of column j, jump selection for every contig (MCA:279-331), per-column constants (the CTA must synchronise
// before using them).  s_cc[(j-1)&1].pc.B holds the base of column j-1.
// One inter-contig jump source: column best (score without the jump cost), its length, layout position.
struct PkSrc { int32_t s; uint32_t l, b; };
__device__ __forceinline__ bool pk_src_before(const PkSrc &x, const PkSrc &y) {   // x is preferred to y: (score, length, position) lexicographic
    return x.s != y.s ? x.s > y.s : (x.l != y.l ? x.l > y.l : x.b > y.b);
}
template <int W>
__device__ __forceinline__ void pk_select_consts(const PackCtx &X, PackSmem &S, ColRec *colrec, int32_t *gcol, const uint8_t *read,
                                                 uint32_t j, PkColConst *s_cc, bool writer, bool ck_col) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, par = j & 1u, C = X.C;
    const int32_t Bprev = s_cc[par ^ 1u].pc.B;
    if (tid < ((C + 31u) & ~31u)) {   // whole warps: one thread per contig
        const uint8_t q = read[j - 1];
        // The three best sources over ALL contigs in the order select_jump's scan prefers (it keeps the last of the longest of
        // the best-scoring candidates: the lexicographic maximum of (score, length, position)); the inter-contig source of a
        // contig is the first of them that is neither the contig nor its opposite strand.  Per lane a sorted top 3 of its own
        // contigs, then three warp-wide pops (REDUX); every participating warp computes the same list.
        constexpr uint32_t NONE = 0xffffffffu;
        PkSrc t0{INT32_MIN, 0u, NONE}, t1 = t0, t2 = t0;
        for (uint32_t b = lane; b < C; b += 32) {
            PkSrc c{S.cm[b], S.cml[b], b};
            if (pk_src_before(c, t0)) { t2 = t1; t1 = t0; t0 = c; }
            else if (pk_src_before(c, t1)) { t2 = t1; t1 = c; }
            else if (pk_src_before(c, t2)) t2 = c;
        }
        PkSrc top[3];
        STITCH_UNROLL
        for (int r = 0; r < 3; ++r) {
            const int32_t ms = __reduce_max_sync(FULL, t0.s);
            const bool c1 = t0.s == ms && t0.b != NONE;
            const uint32_t ml = __reduce_max_sync(FULL, c1 ? t0.l : 0u);
            const bool c2 = c1 && t0.l == ml;
            const int32_t mb = __reduce_max_sync(FULL, c2 ? (int32_t)t0.b : -1);
            top[r].s = ms; top[r].l = ml; top[r].b = mb < 0 ? NONE : (uint32_t)mb;
            if (c2 && (int32_t)t0.b == mb) { t0 = t1; t1 = t2; t2.s = INT32_MIN; t2.l = 0u; t2.b = NONE; }
        }
        const int32_t g = top[0].s;   // best cell of column j-1 over all contigs: the base of column j
        const PkColConst ccl = pk_col_const(X.pk, X.sc, g, Bprev, j, X.n, q);
        const uint32_t a = tid;
        if (a < C) {
            // MCA:279-331, as select_jump (dp_core.h) with the scan over the other contigs replaced by the list above
            JumpInfo J;
            J.score = S.cm[a] + X.sc.g_same; J.len = S.cml[a] + 1; J.idx = X.ent[a].contig_idx; J.from = S.cmk[a];
            const int32_t opp = X.ent[a].opp;
            if (opp >= 0) {
                const int32_t so = S.cm[opp] + X.sc.g_opp;
                if (so > J.score) { J.score = so; J.len = S.cml[opp] + 1; J.idx = X.ent[opp].contig_idx; J.from = S.cmk[opp]; }
            }
            PkSrc it = top[0];
            if (it.b == a || (int32_t)it.b == opp) { it = top[1]; if (it.b == a || (int32_t)it.b == opp) it = top[2]; }
            if (it.b != NONE && it.s + X.sc.g_inter > J.score) {
                J.score = it.s + X.sc.g_inter; J.len = it.l + 1; J.idx = X.ent[it.b].contig_idx; J.from = S.cmk[it.b];
            }
            if (writer) {
                ColRec cr; cr.jscore = J.score; cr.jlen = J.len; cr.jidx = J.idx; cr.jfrom = J.from;
                cr.lx = 0; cr.pad0 = cr.pad1 = cr.pad2 = 0;
                colrec[(uint64_t)j * C + a] = cr;
            }
            S.Jc[a] = pk_jc(X.pk, ccl.pc, J.score, J.len);
            if (X.quiet) pk_quiet_contig(X, S, a, j, ccl, !ck_col);   // (every tile is computed and stored in a checkpoint column)
        }
        if (tid == 0) {
            if (writer) gcol[j - 1] = g;
            s_cc[par] = ccl;
            STITCH_UNROLL
            for (int k = 0; k < PKQ_L + 2; ++k) s_cc[par].yq[k] = j >= 1u + (uint32_t)k ? read[j - 1 - (uint32_t)k] : (uint8_t)0;
        }
    }
}

// After the bulk pass: best score of column n, and the column the tail restarts from: the last checkpoint before
// the first column that can hold a final y-suffix tracker (dp_core.h: first_candidate_column); column n is always
// part of the tail.  Also the score below which a cell cannot hold a final tracker value.
template <int W>
__device__ uint32_t pk_tail_start(const Params &P, PackSmem &S, const Scoring &sc, int32_t *gcol, uint32_t n, uint32_t C, uint32_t K,
                                  int32_t *s_gmax, uint32_t *s_first, int32_t &track_thr, bool writer, bool best_only) {
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t T = W * 32;
    __syncthreads();
    if (tid == 0) {
        int32_t g = S.cm[0];
        for (uint32_t a = 1; a < C; ++a) g = S.cm[a] > g ? S.cm[a] : g;
        if (writer) gcol[n] = g;
        *s_gmax = g; *s_first = n;
    }
    __syncthreads();
    track_thr = MIN_SCORE;
    if (P.tracked_mode) {
        int32_t g = *s_gmax;
        for (uint32_t jj = tid; jj < n; jj += T) g = gcol[jj] > g ? gcol[jj] : g;
        atomicMax(s_gmax, g);
        __syncthreads();
        const int32_t gmax = *s_gmax;
        const int32_t thr = gmax - track_margin(sc, best_only);   // (dp_core.h: reads walked from the best end only need the cells that hold the best score)
        track_thr = thr;
        uint32_t first = n;
        for (uint32_t jj = 1 + tid; jj < n; jj += T) if (gcol[jj] >= thr) { first = jj; break; }
        atomicMin(s_first, first);
        __syncthreads();
    }
    uint32_t j0 = ((*s_first - 1) / K) * K;
    for (uint32_t k = 1; k <= PK_LATE_CKS; ++k) {   // the latest late checkpoint that is still before the first candidate column
        if (n <= PK_LATE_STEP * k) break;
        const uint32_t jl = n - PK_LATE_STEP * k;
        if (jl % K == 0u) continue;   // (a regular checkpoint column: j0 already covers it)
        if (jl < *s_first) { if (jl > j0) j0 = jl; break; }
    }
    __syncthreads();
    return j0;
}

template <int W> __device__ __noinline__ void pk_walk_phase(const Params P, unsigned char *smem_raw);   // kernels_walk.cuh

// ---------------------------------------------------------------------------------------------
// bulk fill
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(W * 32, 1) fill_packed_kernel(const Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PackSmem S; S.carve(smem_raw, P.pso, P.cmax);
    __shared__ uint32_t sJob;
    __shared__ PkColConst s_cc[2];
    __shared__ int32_t s_gmax;
    __shared__ uint32_t s_first;
    __shared__ uint32_t s_cta_lo[17];
    __shared__ PkColStat s_cs;
    const uint32_t tid = threadIdx.x;
    const Scoring sc = P.sc;

    Team team; team.rank = 0; team.size = P.cluster_size;
    if (team.size > 1) team.rank = cg::this_cluster().block_rank();
    if (tid < (uint32_t)W * 4u) mbar_init(smem_u32(S.mbar + tid), 1u);   // staging ring: one mbarrier per (warp, slot)
    if (tid < (uint32_t)W) S.mphase[tid] = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint32_t team_id = blockIdx.x / team.size;

    for (;;) {
        team.sync();
        if (team.rank == 0 && tid == 0) {
            const uint32_t jn = atomicAdd(P.counter, 1u);
            for (uint32_t r = 0; r < team.size; ++r) *team.peer(&sJob, r) = jn;
        }
        team.sync();
        if (sJob >= P.n_jobs) break;
        const uint32_t job = P.order[sJob];
        const JobDesc jd = P.jobs[job];
        const LayoutDesc ld = P.layouts[jd.layout];
        const uint32_t C = ld.C, PM = ld.PM, n = jd.n, K = P.K;
        PackCtx X;
        X.team = team;
        X.pk = P.pk; X.sc = sc;   // (jd.LB is the launch's LB for every packed job)
        // the layout's contig and tile-owner tables are staged in shared memory (every CTA of a team holds its own copy)
        for (uint32_t a = tid; a < C; a += W * 32) S.ent_s[a] = P.ents[ld.ent_off + a];
        for (uint32_t t = tid; t < ld.n_tiles; t += W * 32) S.owner_s[t] = P.owners[ld.owner_off + t];
        X.ent = S.ent_s; X.owner = S.owner_s;
        X.C = C; X.NT = ld.n_tiles; X.bases = P.contig_bases;
        X.n = n; X.yclip_mode = sc.yp != MIN_SCORE && sc.xp == MIN_SCORE;
        X.win_lo = 0; X.skip_below = 0; X.no_finish = false;
        pk_set_ownership(X, W);
        if (tid <= team.size) {
            const uint32_t GW = team.size * W, Weff = X.NT < GW ? X.NT : GW;
            s_cta_lo[tid] = pk_chunk_lo(X.NT, Weff, tid * W);
        }
        X.cta_lo = s_cta_lo;
        X.cluster_smem = P.cluster_state_smem != 0;
        X.cstate = reinterpret_cast<int32_t *>(S.stage);
        if (X.cluster_smem) {   // the rolling state lives in the cluster's shared memory: no HBM traffic per column
            X.Sst = X.cstate - (size_t)X.own_lo * ST; X.state_smem = true; X.staged = false; X.stage_depth = 2;
        } else {
            X.Sst = P.pstate + (uint64_t)team_id * P.pstate_stride; X.state_smem = false; X.staged = true;
            X.stage_depth = P.stage_depth;
        }
        X.tbases = nullptr;
        if (X.staged) {   // the bases of this CTA's tiles in tile order (one bulk copy per tile whatever its contig)
            uint8_t *tbw = P.ptbases + (uint64_t)team_id * P.ptbases_stride + PackSmem::STAGE_PRE;
            X.tbases = tbw;
            const uint32_t lane = tid & 31u;
            __syncthreads();   // S.ent_s / S.owner_s
            for (uint32_t t = X.own_lo + (tid >> 5); t < X.own_hi; t += W) {
                const ContigEntry e = X.ent[X.owner[t]];
                // (a strip starts at a multiple of 8 and contigs are 16-byte aligned: over-reads stay inside the blob's padding)
                *reinterpret_cast<uint2 *>(tbw + (size_t)t * TILE + lane * STRIP) =
                    *reinterpret_cast<const uint2 *>(X.bases + e.seq_off + (t - e.tile_start) * TILE + lane * STRIP);
            }
            fence_async_proxy();
        }
        X.Dst = X.Sst + TILE;
        ColRec *colrec = P.colrec + jd.colrec_off;
        int32_t *gcol = P.gcol + jd.gcol_off;
        const uint8_t *read = P.reads + jd.read_off;
        PkColOut O; O.tb_col = nullptr; O.colrec_col = nullptr; O.sn = nullptr; O.last = nullptr; O.track = false; O.lastcol = false;
        O.track_thr = MIN_SCORE;

        const long long t_bulk0 = clock64();
        pk_state_init0<W>(X, S);
        pk_init_halos<W>(X, S, 0);
        if (tid == 0) { s_cc[0].pc.B = 0; s_cc[0].pc.delta = 0; }
        X.quiet = P.quiet != 0 && team.size == 1 && !X.cluster_smem;
        X.quiet_first = X.quiet && P.quiet_first != 0; X.quiet_edge = X.quiet && P.quiet_edge != 0;
        X.quiet_last = X.quiet && P.quiet_last != 0;
        if (tid < 2 * W) S.haloF[tid] = 0;
        if (tid == 0) { s_cs.skipped = 0; s_cs.timing = P.dbg ? 1u : 0u; s_cs.t_tiles = s_cs.t_finish = s_cs.t_busy = s_cs.t_select = s_cs.t_f1 = s_cs.t_f2 = s_cs.t_fa = 0; for (int w = 0; w < 32; ++w) { s_cs.t_warp[w] = 0; s_cs.n_warp[w] = 0; } }
        if (X.quiet) {   // quiet tiles: no tile is quiet yet; base classes of every tile
            for (uint32_t a = tid; a < C; a += W * 32) S.Q[a] = pk_quiet_init(X.pk);
            const uint32_t lane = tid & 31u;
            for (uint32_t t = tid >> 5; t < X.NT; t += W) {
                const ContigEntry e = X.ent[X.owner[t]];
                const uint32_t r0 = (t - e.tile_start) * TILE + lane * STRIP;
                uint8_t xq[STRIP];
                unpack8(*reinterpret_cast<const uint2 *>(X.bases + e.seq_off + r0), xq);
                uint32_t bits = 0;
                STITCH_UNROLL
                for (int k = 0; k < STRIP; ++k) if (r0 + (uint32_t)k + 1 < e.m) bits |= pk_base_bit(xq[k]);   // ordinary rows (i < m) only
                bits = __reduce_or_sync(FULL, bits);
                const uint32_t tic = t - e.tile_start;
                if (tic + 1 != e.ntiles || (X.quiet_last && tic != 0)) bits |= PackSmem::TB_CANQ;
                if (lane == 0) S.tb[t] = (uint8_t)bits;
            }
        }
        __syncthreads();

        for (uint32_t j = 1; j <= n; ++j) {
            const uint32_t par = j & 1u;
            const long long cs0 = clock64();
            const uint32_t late = pk_late_ck(j, n, K);
            const bool ck_col = ((j % K == 0) && j < n) || late != 0u;
            pk_select_consts<W>(X, S, colrec, gcol, read, j, s_cc, team.rank == 0, ck_col);
            __syncthreads();
            const PkColConst cc = s_cc[par];
            if (tid == 0 && P.dbg) s_cs.t_select += (unsigned long long)(clock64() - cs0);
            if (X.quiet) pk_column<W, false, true>(X, S, cc.pc, cc.r0pkey, cc.cr1key, j, O, cc.yq, &s_cs);
            else pk_column<W, false>(X, S, cc.pc, cc.r0pkey, cc.cr1key, j, O);
            if (ck_col) {
                const uint64_t ci = pk_ck_index(j, n, K);
                pk_write_ck<W>(X, S, P.pck + jd.ck_off + ci * 2 * PM, P.ck_sum + jd.cksum_off + ci * C);
            }
        }
        if (tid == 0 && team.rank == 0 && P.qstats) {
            atomicAdd(P.qstats + 0, (unsigned long long)X.NT * n);
            if (s_cs.skipped) atomicAdd(P.qstats + 1, (unsigned long long)s_cs.skipped);
            if (P.dbg) {
                atomicAdd(P.dbg + 7, s_cs.t_select); atomicAdd(P.dbg + 8, s_cs.t_tiles); atomicAdd(P.dbg + 9, s_cs.t_finish);
                atomicAdd(P.dbg + 10, s_cs.t_busy / W); atomicAdd(P.dbg + 11, s_cs.t_f1); atomicAdd(P.dbg + 12, s_cs.t_f2); atomicAdd(P.dbg + 13, s_cs.t_fa);
                for (int w = 0; w < W; ++w) { atomicAdd(P.dbg + 16 + w, s_cs.t_warp[w]); atomicAdd(P.dbg + 48 + w, s_cs.n_warp[w]); }
            }
        }
        int32_t track_thr = MIN_SCORE;
        if (team.size > 1) {   // gcol / colrec of the whole read (written by rank 0) must be visible to the team
            if (tid == 0 && team.rank == 0) { int32_t g = S.cm[0]; for (uint32_t a = 1; a < C; ++a) g = S.cm[a] > g ? S.cm[a] : g; gcol[n] = g; }
            team.sync();
        }
        const uint32_t j0 = pk_tail_start<W>(P, S, sc, gcol, n, C, K, &s_gmax, &s_first, track_thr, team.rank == 0, jd.walk == host::WALK_BEST);
        if (tid == 0 && team.rank == 0) P.tail_j0[job] = j0;
        __syncthreads();
        const long long t_tail0 = clock64();
        pk_tail<W>(P, jd, ld, X, S, &s_cc[0], j0, track_thr);
        if (P.dbg && tid == 0 && team.rank == 0) {
            atomicAdd(P.dbg + 0, (unsigned long long)(n - j0));
            atomicAdd(P.dbg + 1, (unsigned long long)(clock64() - t_tail0));
            atomicAdd(P.dbg + 2, (unsigned long long)(t_tail0 - t_bulk0));
        }
        if (P.done) {   // publish: every record of this read is written
            __syncthreads();
            if (tid == 0) { __threadfence(); atomicExch(P.done + job, 1u); }
        }
    }
    // ---- second phase of the persistent kernel (single-CTA teams): CTAs that find the fill queue empty walk the
    // reads in the order they were filled, so the end-of-read fix-ups and walks fill the SMs that the last fills
    // leave idle.  A CTA only waits for reads that other, running CTAs of this launch are still filling. ----
    if (P.done) pk_walk_phase<W>(P, smem_raw);
}

// ---------------------------------------------------------------------------------------------
// walk: packed re-fill of one unit (contig `a`, the block of K columns holding column j)
// ---------------------------------------------------------------------------------------------
// Per-unit staging in shared memory: the jump records, column bases and read bases of the unit's columns and
// the contig's bases, so that a re-filled column touches global memory only for its traceback bytes.
struct UnitStage {
    JumpInfo *J;        // [K]
    int32_t *B;         // [K + 1]: B[t] = base of column jb + t  (B[0] = base of the checkpointed column jb)
    uint8_t *q;         // [K]
    uint8_t *bases;     // contig bases (+ 16 bytes of padding for the strip over-read); only when Params::unit_stage_bases
    PkColConst *cc;     // [K] cone re-fills: the per-column constants of the whole unit (nullptr: not carved, no cone re-fills)
    int32_t *Jc;        // [K]
    static __host__ __device__ size_t r16(size_t v) { return (v + 15) / 16 * 16; }
    static bool cone_fits(uint32_t K) { return K <= PK_CONE_MAX_COLS; }
    // `stage_bases` = false: the longest contig does not fit beside the rest; the re-fill reads bases from global memory.
    // The cone area is carved whenever the checkpoint spacing allows cone re-fills at all (cone_fits).
    static size_t bytes(uint32_t K, uint32_t max_ctiles, bool stage_bases) {
        return r16(sizeof(JumpInfo) * K) + r16(sizeof(int32_t) * (K + 1)) + r16(K) + (cone_fits(K) ? r16(sizeof(PkColConst) * K) + r16(sizeof(int32_t) * K) : 0) +
               (stage_bases ? r16((size_t)max_ctiles * TILE + 16) : 0);
    }
    __device__ void carve(unsigned char *raw, uint32_t K) {   // raw is 16-byte aligned; every part stays 16-byte aligned
        J = reinterpret_cast<JumpInfo *>(raw);
        B = reinterpret_cast<int32_t *>(raw + r16(sizeof(JumpInfo) * K));
        q = reinterpret_cast<uint8_t *>(B) + r16(sizeof(int32_t) * (K + 1));
        unsigned char *p = q + r16(K);
        cc = nullptr; Jc = nullptr;
        if (K <= PK_CONE_MAX_COLS) { cc = reinterpret_cast<PkColConst *>(p); p += r16(sizeof(PkColConst) * K); Jc = reinterpret_cast<int32_t *>(p); p += r16(sizeof(int32_t) * K); }
        bases = p;
    }
};

template <int W>
__device__ void pk_refill_unit(const Params &P, const JobDesc &jd, const LayoutDesc &ld, PackSmem &S, UnitStage &U, ContigEntry *s_en,
                               PkColConst *s_cc, uint32_t a, uint32_t j, uint32_t i_entry, int32_t *pstate, uint64_t pstate_half, bool state_smem,
                               uint8_t *bytes, ColRec *ucr, TbUnit *unit_out) {
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t T = W * 32;
    const uint32_t C = ld.C, PM = ld.PM, n = jd.n, K = P.K;
    const uint32_t b = (j - 1) / K, jb = b * K;
    const uint32_t ncols = j - jb;   // columns (jb, j]: the walk enters at column j and only moves left
    const ContigEntry gen = P.ents[ld.ent_off + a];
    const uint32_t pm = gen.ntiles * TILE, gbase = gen.tile_start * TILE;
    const int32_t *gcol = P.gcol + jd.gcol_off;
    const ColRec *colrec = P.colrec + jd.colrec_off;
    const uint8_t *read = P.reads + jd.read_off;
    const bool stage_bases = P.unit_stage_bases != 0;
    // cone of the entry cell (dp_packed.h): the columns run over the window of tiles the cone ever touches, split into one
    // fixed chunk per warp; a chunk that lies entirely above the cone's top in a column is left stale.  Short units also get
    // every per-column constant precomputed (one barrier per column).
    PkCone cone; cone.on = false; cone.slope = 0; cone.win_lo = 0; cone.win_n = gen.ntiles;
    if (P.cone) cone = pk_cone_plan(P.sc, gen, i_entry, j, jb);
    const bool pre = cone.on && U.cc != nullptr && ncols <= PK_CONE_MAX_COLS;
    const PK pk = P.pk;
    if (tid == 0) { *s_en = gen; s_en->tile_start = 0; if (stage_bases) s_en->seq_off = 0; }
    for (uint32_t t = tid; t < ncols; t += T) {
        const ColRec cr = colrec[(uint64_t)(jb + 1 + t) * C + a];
        JumpInfo J; J.score = cr.jscore; J.len = cr.jlen; J.idx = cr.jidx; J.from = cr.jfrom;
        U.J[t] = J;
        U.q[t] = read[jb + t];
    }
    for (uint32_t t = tid; t <= ncols; t += T) U.B[t] = (jb + t >= 1) ? gcol[jb + t - 1] : 0;
    if (stage_bases) {
        const uint32_t lo = cone.on ? cone.win_lo * TILE : 0u, hi = cone.on ? (cone.win_lo + cone.win_n) * TILE + 16u : pm + 16u;
        for (uint32_t t = (lo >= 16u ? lo - 16u : 0u) + tid; t < hi; t += T) U.bases[t] = t < gen.m ? P.contig_bases[gen.seq_off + t] : (uint8_t)0;
    }
    __syncthreads();
    PackCtx X;
    X.pk = pk; X.sc = P.sc; X.ent = s_en; X.owner = nullptr; X.C = 1; X.bases = stage_bases ? U.bases : P.contig_bases;
    X.Sst = pstate; X.Dst = pstate + TILE; X.n = n; X.yclip_mode = P.sc.yp != MIN_SCORE && P.sc.xp == MIN_SCORE;
    X.team.rank = 0; X.team.size = 1; X.state_smem = state_smem; X.staged = false; X.stage_depth = 2; X.tbases = nullptr;   // bases are staged in shared memory here
    X.NT = cone.on ? cone.win_n : gen.ntiles; X.win_lo = cone.on ? cone.win_lo : 0u; X.skip_below = 0; X.no_finish = cone.on;
    pk_set_ownership(X, W); X.cluster_smem = false; X.quiet = false; X.quiet_first = false; X.quiet_edge = false; X.quiet_last = false; X.cstate = nullptr; X.cta_lo = nullptr;
    if (b == 0) pk_state_init0<W>(X, S);
    else pk_state_from_ck<W>(X, S, P.pck + jd.ck_off + (uint64_t)(b - 1) * 2 * PM + 2 * gbase, P.ck_sum + jd.cksum_off + (uint64_t)(b - 1) * C + a,
                             U.B[0]);
    pk_init_halos<W>(X, S, jb & 1u);
    PkColOut O; O.sn = nullptr; O.last = nullptr; O.track = false; O.lastcol = false; O.track_thr = MIN_SCORE;
    // the rows above the window are stale by construction: the halo of the window's first tile is any in-band value
    if (cone.on && cone.win_lo > 0 && tid < 17u)
        for (uint32_t par = 0; par < 2; ++par) {
            if (tid < 9u) S.haloS[(par * W) * 9 + tid] = pk.NEGKEY;
            else S.haloD[(par * W) * 8 + tid - 9u] = pk.NEGKEY + pk.PD6;
        }
    if (pre) {
        // every per-column constant of the unit at once (one column per thread), so that a column is tiles + ONE barrier
        for (uint32_t t = tid; t < ncols; t += T) {
            const int32_t B = U.B[t + 1], Bprev = U.B[t];   // B[t+1] = G(jj-1) = base of column jj
            PCol pcl; pcl.B = B; pcl.delta = B - Bprev;
            U.Jc[t] = pk_jc(pk, pcl, U.J[t].score, U.J[t].len);
            U.cc[t] = pk_col_const(pk, P.sc, B, Bprev, jb + 1 + t, n, U.q[t]);
        }
        __syncthreads();
        PackSmem S2 = S;
        for (uint32_t jj = jb + 1; jj <= j; ++jj) {
            const uint32_t t = jj - jb - 1;
            const PkColConst cc = U.cc[t];
            S2.Jc = U.Jc + t;
            X.skip_below = pk_cone_top_tile(cone, i_entry, j, jj);
            O.tb_col = bytes + (uint64_t)t * pm; O.colrec_col = nullptr;
            pk_column<W, true>(X, S2, cc.pc, cc.r0pkey, cc.cr1key, jj, O);   // (no per-contig finish: ends after its first barrier)
        }
    } else {
        __syncthreads();
        for (uint32_t jj = jb + 1; jj <= j; ++jj) {
            const uint32_t t = jj - jb - 1;
            if (tid == 0) {
                const int32_t B = U.B[t + 1], Bprev = U.B[t];
                const JumpInfo J = U.J[t];
                PCol pcl; pcl.B = B; pcl.delta = B - Bprev;
                S.Jw[0] = J;
                S.Jc[0] = pk_jc(pk, pcl, J.score, J.len);
                *s_cc = pk_col_const(pk, P.sc, B, Bprev, jj, n, U.q[t]);
            }
            __syncthreads();
            const PkColConst cc = *s_cc;
            if (cone.on) X.skip_below = pk_cone_top_tile(cone, i_entry, j, jj);
            O.tb_col = bytes + (uint64_t)t * pm; O.colrec_col = cone.on ? nullptr : ucr + t;
            pk_column<W, true>(X, S, cc.pc, cc.r0pkey, cc.cr1key, jj, O);
        }
    }
    if (tid == 0) {
        unit_out->bytes = bytes; unit_out->cr = ucr; unit_out->a = a; unit_out->jb = jb; unit_out->je = j; unit_out->pm = pm;
        unit_out->i_hi = cone.on ? i_entry : 0xffffffffu; unit_out->slope = cone.slope;
    }
    __syncthreads();
}

}  // namespace gpu
}  // namespace stitch
