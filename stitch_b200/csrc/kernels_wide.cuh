// kernels_wide.cuh — the exact, fully general ("wide": i32 score + u32 length) kernels.
//
//   column_wide     one column of a set of contig-strands (the body of MCA::custom's loop,
//                   multi_contig_aligner.rs:270-347 = SCA:188-239 + 292-451 of the reference);
//                   round-based: W warp tiles per round, warp-shuffle max-plus scan of the insertion
//                   chain, cross-warp carry folding through shared memory.  Valid for every scoring.
//   fill_wide_kernel   columns (j0, n] of a read over all contig-strands, from column 0 or from the
//                   wide state the packed kernel hands over (kernels_packed.cuh); keeps the
//                   y-suffix trackers for the last columns and the column-n records.
//   fixup_kernel    end-of-read fix-up of column n (SCA:453-555) + the check that the tracking
//                   window was wide enough.
//   walk_kernel     end-contig selection + pointer walk (traceback/mod.rs:129-373); the packed
//                   traceback bytes are re-filled on demand, one (contig, block of columns) unit
//                   at a time, from the column-state checkpoints (checkpoint-and-recompute).
#pragma once
#include <cuda_runtime.h>

#include "dp_core.h"
#include "dp_packed.h"
#include "host_common.hpp"

namespace stitch {
namespace gpu {

constexpr unsigned FULL = 0xffffffffu;

struct JobDesc {
    uint64_t read_off;     // into the reads blob
    uint64_t colrec_off;   // ColRec records  [(n+1) * C]
    uint64_t cell_off;     // LastCell / SnRec records [PM]
    uint64_t ck_off;       // CellState records [(nb-1) * PM]
    uint64_t cksum_off;    // CkSum records [(nb-1) * C]
    uint64_t gcol_off;     // int32 [n+1]
    uint64_t ops_off;      // OutOp records
    uint64_t hand_off;     // CellState records [PM]: wide state at column j0 (packed -> wide hand-over)
    uint64_t handsum_off;  // CkSum records [C]
    uint32_t j0, LB;       // columns [1, j0] run on the packed kernel (0: none); LB = length bits of its keys
    uint32_t n, layout;
    uint32_t walk, from_contig;
    uint32_t ops_cap, chain_first, max_chains, track_from;
};

struct LayoutDesc { uint32_t ent_off, C, n_tiles, owner_off, pad0, PM, max_ctiles, pad1; };

enum : uint32_t { JOB_OK = 0, JOB_NEED_FULL_TRACK = 100 };
struct JobOut { uint32_t n_chains, status; };

// Byte offsets of the tables of PackSmem (kernels_packed.cuh) inside the dynamic shared memory of a launch.  The host
// computes them once per launch (PackSmem::layout) so that a table address is one constant-bank value added to the base:
// the packed kernels have no registers to spare for ~20 table pointers, and re-deriving them from cmax / ntmax inside the
// tile loop cost ~40 instructions per tile.
struct PackSmemOff {
    uint32_t Jw, stash, Jc, cm, Sm, SmKey, cml, cmk, slm, tbm, DmKey, tilemax, haloS, haloD, haloF, Q, ent_s, owner_s, clist, tb, mbar, mphase, end;
};

struct Params {
    Scoring sc;
    PK pk;                   // packed kernels: the key layout of the launch (one LB for every packed job: the largest any of them needs)
    PackSmemOff pso;         // packed kernels: table offsets in dynamic shared memory
    const JobDesc *jobs;
    const uint32_t *order;
    uint32_t n_jobs, cmax;
    const LayoutDesc *layouts;
    const ContigEntry *ents;
    const uint16_t *owners;
    const uint8_t *contig_bases;
    const uint8_t *reads;
    CellState *state;        // per CTA: two rolling column buffers
    uint64_t state_stride;   // CellState records per CTA
    uint64_t state_half;     // records per column buffer
    uint8_t *unit_bytes;     // per CTA (walk): packed traceback bytes of the loaded unit
    uint64_t unit_stride;
    ColRec *colrec;
    LastCell *last;
    SnRec *sn;
    CellState *ck_state;     // checkpoints of the wide path
    int32_t *pck;            // checkpoints of the packed path: raw copies of the packed state (2 * PM keys each)
    CkSum *ck_sum;
    CellState *hand_state;
    CkSum *hand_sum;
    int32_t *pstate;         // packed kernel, per CTA: S keys then D keys
    uint64_t pstate_stride, pstate_half;
    uint8_t *ptbases;        // packed kernel, per team: the contig bases of the read's layout in TILE order (tile t at 16 + t * TILE),
    uint64_t ptbases_stride; // so that a tile's bases are one bulk copy away whatever contig it belongs to
    uint32_t ntmax;          // largest tile count among the packed jobs
    uint32_t cluster_size;   // packed kernel: CTAs per read (thread-block cluster), 1 = no cluster
    uint32_t stage_bytes;    // packed kernels: size of the front area of their dynamic shared memory (PackSmem)
    uint32_t stage_depth;    // packed fill: slots per warp of the bulk-copy staging ring (2 or 4)
    uint32_t cluster_state_smem;   // the rolling state lives in the cluster's shared memory (bytes per CTA), 0 = global memory
    uint32_t quiet;          // packed bulk pass: skip quiet tiles (dp_packed.h), single-CTA teams only
    uint32_t quiet_tail;     // ... and in the tail columns (traceback variant), for tiles below the tracking threshold
    uint32_t quiet_first, quiet_edge, quiet_last;   // ... also the first / last tile of a contig, the first and last tile of a warp chunk
    unsigned long long *qstats;   // [0] tile-columns of the bulk passes, [1] of those skipped as quiet
    uint32_t *tail_j0;       // per job: the checkpointed column the packed tail restarts from
    int32_t *wpstate;        // walk kernel, per CTA: packed state of one contig (S keys then D keys)
    uint64_t wpstate_stride, wpstate_half;
    ColRec *unit_cr;         // walk kernel, per CTA: K per-column records of the loaded unit
    uint32_t max_ctiles;     // most tiles of any single contig (walk kernel shared memory)
    unsigned long long *dbg; // optional counters (STITCH_DEBUG_STATS): see cuda_backend.cu
    uint32_t *done;          // packed kernel with the in-kernel walk phase: per job, 1 once its fill and tail are complete
    uint32_t walk_stage_smem_off;   // walk kernel: byte offset of the per-unit staging area in dynamic shared memory
    uint32_t walk_state_smem_off;   // walk kernel: byte offset of the packed unit state in dynamic shared memory (0: global)
    uint32_t cone;                  // walk: packed re-fills restricted to the cone of the entry cell (dp_packed.h); 0: whole contigs
    uint32_t unit_stage_bases;      // walk: the per-unit staging area holds the contig's bases (they fit); 0: read from global memory
    int32_t *gcol;
    OutOp *ops;
    ChainHdr *chains;
    JobOut *job_out;
    uint32_t *counter;
    uint32_t K;              // checkpoint spacing (columns)
    int tracked_mode;        // ys != MIN_SCORE
    int force_full;          // re-run of reads whose tracking window was too narrow
    const uint32_t *redo_j0; // force_full: per job, the checkpointed column the re-run starts from (multiple of K, 0 = column 0)
};

__device__ __forceinline__ ICarry shfl_up_ic(ICarry c, int d) {
    ICarry r;
    r.v = __shfl_up_sync(FULL, c.v, d);
    r.il = __shfl_up_sync(FULL, c.il, d);
    r.open = __shfl_up_sync(FULL, c.open, d);
    return r;
}
__device__ __forceinline__ XsPart shfl_xor_xs(XsPart p, int d) {
    XsPart r;
    r.t = __shfl_xor_sync(FULL, p.t, d); r.len = __shfl_xor_sync(FULL, p.len, d); r.row = __shfl_xor_sync(FULL, p.row, d);
    return r;
}
__device__ __forceinline__ CmPart shfl_xor_cm(CmPart p, int d) {
    CmPart r;
    r.S = __shfl_xor_sync(FULL, p.S, d); r.row = __shfl_xor_sync(FULL, p.row, d);
    r.sl = __shfl_xor_sync(FULL, p.sl, d); r.valid = __shfl_xor_sync(FULL, p.valid, d);
    return r;
}

// Shared-memory working set of column_wide for up to `cmax` contigs.
template <int W>
struct WideSmem {
    JumpInfo *J; RowM *rowm; XsPart *xs; CmPart *cmp;
    int32_t *cm; uint32_t *cml, *cmk; int32_t *Sm; uint32_t *slm, *tbm;
    ICarry *tile_agg;   // [2][W]
    ICarry *round;      // [2]
    static size_t bytes(uint32_t cmax) {
        return sizeof(JumpInfo) * cmax + sizeof(RowM) * cmax + (sizeof(XsPart) + sizeof(CmPart)) * cmax * W +
               sizeof(int32_t) * cmax * 6 + sizeof(ICarry) * (2 * W + 2) + 64;
    }
    __device__ void carve(unsigned char *raw, uint32_t cmax) {
        J = reinterpret_cast<JumpInfo *>(raw);
        rowm = reinterpret_cast<RowM *>(J + cmax);
        xs = reinterpret_cast<XsPart *>(rowm + cmax);
        cmp = reinterpret_cast<CmPart *>(xs + (size_t)cmax * W);
        cm = reinterpret_cast<int32_t *>(cmp + (size_t)cmax * W);
        cml = reinterpret_cast<uint32_t *>(cm + cmax);
        cmk = cml + cmax;
        Sm = reinterpret_cast<int32_t *>(cmk + cmax);
        slm = reinterpret_cast<uint32_t *>(Sm + cmax);
        tbm = slm + cmax;
        tile_agg = reinterpret_cast<ICarry *>(tbm + cmax);
        round = tile_agg + 2 * W;
    }
};

struct ColWide {
    const ContigEntry *ent; const uint16_t *owner;   // owner == nullptr: a single contig (position 0)
    uint32_t C, NT;
    const uint8_t *bases;
    const CellState *prev; CellState *curr; CellState *ck;
    uint8_t *tb_col; ColRec *colrec_col;
    SnRec *sn; LastCell *last;
    uint32_t j, n;
    uint8_t q;
    bool track, lastcol;
};

// One column.  On entry S.J[a] (jump into contig a for this column) and S.Sm/slm/tbm (row-m summary
// of column j-1) are set and visible to the CTA; on exit S.cm/cml/cmk hold the column best of every
// contig, S.Sm/slm/tbm the new row-m summaries, and the CTA is synchronised.
template <int W>
__device__ void column_wide(const Scoring &sc, const ColWide &A, WideSmem<W> &S) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr uint32_t T = W * 32;
    const uint32_t C = A.C, NT = A.NT, j = A.j, n = A.n;
    const Row0 r0 = row0_at(sc, j, n), r0p = row0_at(sc, j - 1, n);
    ColConst cc; cc.j = j; cc.n = n; cc.q = A.q;
    { const int32_t dj = sc.o + sc.e * (int32_t)j; cc.xclip_score = sc.xp + (sc.yp > dj ? sc.yp : dj); }
    cc.sl0j = r0.sl;
    for (uint32_t x = tid; x < C * W; x += T) { xs_init(S.xs[x]); cm_init(S.cmp[x]); }
    __syncthreads();

    uint32_t par = 0;
    for (uint32_t t0 = 0; t0 < NT; t0 += W, par ^= 1u) {
        const uint32_t tile = t0 + warp;
        const bool tact = tile < NT;
        TileCtx tc; LaneA la; ICarry excl; uint32_t row0 = 0, a = 0; bool lact = false;
        uint8_t x[STRIP];
        excl.v = MIN_SCORE; excl.il = 0; excl.open = 0;
        la.agg = excl; la.has_m = 0;
        if (tact) {
            a = A.owner ? A.owner[tile] : 0u;
            const ContigEntry en = A.ent[a];
            tc.a = a; tc.self_idx = en.contig_idx; tc.m = en.m; tc.tile = tile; tc.tile_in_contig = tile - en.tile_start;
            tc.J = S.J[a]; tc.circular = en.circular != 0; tc.wrap_src_ok = S.tbm[a] != TB_XCLIP_SUFFIX;
            tc.Sm_prev = S.Sm[a]; tc.slm_prev = S.slm[a];
            row0 = tc.tile_in_contig * TILE + lane * STRIP + 1;
            lact = row0 <= en.m;
            CellState up[STRIP];
            STITCH_UNROLL
            for (int k = 0; k < STRIP; ++k) {
                if (lact) {
                    const int4 v = *reinterpret_cast<const int4 *>(A.prev + state_index(tile, lane, (uint32_t)k));
                    up[k].S = v.x; up[k].D = v.y; up[k].sl = (uint32_t)v.z; up[k].dl = (uint32_t)v.w;
                } else { up[k].S = MIN_SCORE; up[k].D = MIN_SCORE; up[k].sl = 0; up[k].dl = 0; }
            }
            if (lact) {
                const uint8_t *xb = A.bases + en.seq_off + row0 - 1;
                STITCH_UNROLL
                for (int k = 0; k < STRIP; ++k) x[k] = (row0 + (uint32_t)k <= en.m) ? xb[k] : (uint8_t)0;
            } else {
                STITCH_UNROLL
                for (int k = 0; k < STRIP; ++k) x[k] = 0;
            }
            int32_t dgS = __shfl_up_sync(FULL, up[STRIP - 1].S, 1);
            uint32_t dgsl = __shfl_up_sync(FULL, up[STRIP - 1].sl, 1);
            if (lane == 0) {
                if (tc.tile_in_contig == 0) { dgS = r0p.S; dgsl = r0p.sl; }
                else {
                    const CellState c = A.prev[state_index(tile - 1, 31, STRIP - 1)];
                    dgS = c.S; dgsl = c.sl;
                }
            }
            if (lact) lane_pass_a(sc, cc, tc, row0, up, dgS, dgsl, x, la, &S.rowm[a]);
            // inclusive max-plus scan of the lane aggregates
            ICarry inc = la.agg;
            STITCH_UNROLL
            for (int d = 1; d < 32; d <<= 1) {
                const ICarry o = shfl_up_ic(inc, d);
                if ((int)lane >= d) inc = icarry_combine(o, (uint32_t)(d * STRIP), sc.e, inc);
            }
            excl = shfl_up_ic(inc, 1);
            if (lane == 31) S.tile_agg[par * W + warp] = inc;
        }
        __syncthreads();
        if (tact) {
            // carry into this tile: fold the aggregates of the tiles of the same contig before it
            uint32_t w0 = warp;
            if (A.owner) { while (w0 > 0 && A.owner[t0 + w0 - 1] == a) --w0; } else w0 = 0;
            ICarry c;
            if (t0 + w0 == A.ent[a].tile_start) c = icarry_row1(sc, r0);
            else c = S.round[par];
            for (uint32_t u = w0; u < warp; ++u) c = icarry_combine(c, TILE, sc.e, S.tile_agg[par * W + u]);
            const uint32_t last_tile = (t0 + W < NT ? t0 + W : NT) - 1;
            if (tile == last_tile && lane == 0) S.round[par ^ 1u] = icarry_combine(c, TILE, sc.e, S.tile_agg[par * W + warp]);
            const ICarry cin = lane == 0 ? c : icarry_combine(c, lane * STRIP, sc.e, excl);
            LaneB lb; xs_init(lb.xs); cm_init(lb.cm);
            if (lact) lane_pass_b(sc, cc, tc, row0, lane, la, cin, A.curr, A.ck, A.tb_col, A.track, A.sn, A.lastcol, A.last, x, lb,
                                  &S.rowm[a]);
            STITCH_UNROLL
            for (int d = 16; d >= 1; d >>= 1) {
                lb.xs = xs_merge(lb.xs, shfl_xor_xs(lb.xs, d));
                lb.cm = cm_merge(lb.cm, shfl_xor_cm(lb.cm, d));
            }
            if (lane == 0) {
                const uint32_t slot = a * W + warp;
                S.xs[slot] = xs_merge(S.xs[slot], lb.xs);
                S.cmp[slot] = cm_merge(S.cmp[slot], lb.cm);
            }
        }
    }
    __syncthreads();

    // per contig: finish row m, column best for the next jump
    for (uint32_t a = warp; a < C; a += W) {
        XsPart xs; CmPart cm; xs_init(xs); cm_init(cm);
        if (lane < (uint32_t)W) { xs = S.xs[a * W + lane]; cm = S.cmp[a * W + lane]; }
        STITCH_UNROLL
        for (int d = 16; d >= 1; d >>= 1) {
            xs = xs_merge(xs, shfl_xor_xs(xs, d));
            cm = cm_merge(cm, shfl_xor_cm(cm, d));
        }
        if (lane == 0) {
            const ContigColOut o = contig_finalize(sc, cc, A.ent[a], a, C, S.rowm[a], xs, cm, r0, S.J[a], A.curr, A.ck, A.tb_col,
                                                   A.colrec_col, A.track, A.sn, A.lastcol, A.last);
            S.cm[a] = o.cm.S; S.cmk[a] = o.cm.row; S.cml[a] = o.cm.sl;
            S.Sm[a] = o.Sm; S.slm[a] = o.slm; S.tbm[a] = o.s_tb_m;
        }
    }
    __syncthreads();
}

// Column 0 of every contig of the layout into `st` (SCA:97-186), row-m summaries and column best.
template <int W>
__device__ void column0_wide(const Scoring &sc, const ContigEntry *ent, const uint16_t *owner, uint32_t C, uint32_t NT,
                             CellState *st, WideSmem<W> &S, bool init_sn, SnRec *sn, uint32_t n, bool write_state = true) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr uint32_t T = W * 32;
    for (uint32_t tile = warp; tile < NT; tile += W) {
        const uint32_t a = owner ? owner[tile] : 0u;
        const ContigEntry en = ent[a];
        const uint32_t tic = tile - en.tile_start;
        STITCH_UNROLL
        for (int k = 0; k < STRIP; ++k) {
            const uint32_t i = tic * TILE + lane * STRIP + (uint32_t)k + 1;
            if (i <= en.m) {
                const Col0 c0 = col0_at(sc, i, en.m);
                CellState s; s.S = c0.S; s.D = MIN_SCORE; s.sl = c0.sl; s.dl = 0;
                const uint32_t si = state_index(tile, lane, (uint32_t)k);
                if (write_state) st[si] = s;
                if (init_sn) sn[si] = sn_init(sc, c0.S, c0.sl, en.contig_idx, n);
            }
        }
    }
    for (uint32_t a = tid; a < C; a += T) {
        const ContigEntry en = ent[a];
        S.cm[a] = 0; S.cml[a] = 0; S.cmk[a] = 0;   // column-0 best is S(0,0) = 0 at row 0
        const Col0 cm = col0_at(sc, en.m, en.m);
        S.Sm[a] = cm.S; S.slm[a] = cm.sl; S.tbm[a] = cm.s_tb;
    }
}

// ---------------------------------------------------------------------------------------------
// fill: persistent CTAs pull reads from a queue
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(W * 32) fill_wide_kernel(const Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WideSmem<W> S; S.carve(smem_raw, P.cmax);
    __shared__ uint32_t sJob;
    const uint32_t tid = threadIdx.x;
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
        const uint32_t C = ld.C, NT = ld.n_tiles, PM = ld.PM, n = jd.n, K = P.K;
        CellState *st0 = P.state + (uint64_t)blockIdx.x * P.state_stride;
        CellState *st1 = st0 + P.state_half;
        ColRec *colrec = P.colrec + jd.colrec_off;
        LastCell *last = P.last + jd.cell_off;
        SnRec *sn = P.sn + jd.cell_off;
        int32_t *gcol = P.gcol + jd.gcol_off;
        const uint8_t *read = P.reads + jd.read_off;
        // columns [1, j0] are already filled: by the packed kernel (hand-over state) or, in a re-run, up to a checkpoint
        const uint32_t j0 = P.force_full ? P.redo_j0[P.order[sJob]] : jd.j0;
        const uint32_t track_from = P.tracked_mode ? (P.force_full ? 1u : (jd.track_from > j0 ? jd.track_from : j0 + 1)) : n + 1;

        column0_wide<W>(sc, ent, owner, C, NT, st0, S, track_from <= n, sn, n, j0 == 0);
        for (uint32_t a = tid; a < C; a += T) {
            int32_t t; uint32_t lx; col0_tracker(sc, ent[a].m, t, lx);
            ColRec cr; cr.jscore = 0; cr.jlen = 0; cr.jidx = 0; cr.jfrom = 0; cr.lx = lx; cr.pad0 = cr.pad1 = cr.pad2 = 0;
            colrec[a] = cr;
        }
        if (j0 > 0) {   // hand-over: wide state and row-m summaries of column j0
            __syncthreads();
            const CellState *hs = P.force_full ? P.ck_state + jd.ck_off + (uint64_t)(j0 / K - 1) * PM : P.hand_state + jd.hand_off;
            const CkSum *hsum = P.force_full ? P.ck_sum + jd.cksum_off + (uint64_t)(j0 / K - 1) * C : P.hand_sum + jd.handsum_off;
            CellState *dst = (j0 & 1u) ? st1 : st0;
            for (uint32_t p = tid; p < PM; p += T) dst[p] = hs[p];
            for (uint32_t a = tid; a < C; a += T) {
                const CkSum cs = hsum[a];
                S.Sm[a] = cs.Sm; S.slm[a] = cs.slm; S.tbm[a] = cs.tbm;
            }
        }
        __syncthreads();

        for (uint32_t j = j0 + 1; j <= n; ++j) {
            if (j == j0 + 1 && j0 > 0) {   // the jump of the first wide column was selected by the packed kernel
                for (uint32_t a = tid; a < C; a += T) {
                    const ColRec cr = colrec[(uint64_t)j * C + a];
                    JumpInfo J; J.score = cr.jscore; J.len = cr.jlen; J.idx = cr.jidx; J.from = cr.jfrom;
                    S.J[a] = J;
                }
            } else {
                for (uint32_t a = tid; a < C; a += T) S.J[a] = select_jump(sc, ent, C, a, S.cm, S.cml, S.cmk);
                if (tid == 0) {
                    int32_t g = S.cm[0];
                    for (uint32_t a = 1; a < C; ++a) g = S.cm[a] > g ? S.cm[a] : g;
                    gcol[j - 1] = g;
                }
            }
            const bool ck = (j % K == 0) && j < n;
            ColWide A;
            A.ent = ent; A.owner = owner; A.C = C; A.NT = NT; A.bases = P.contig_bases;
            A.prev = (j & 1u) ? st0 : st1; A.curr = (j & 1u) ? st1 : st0;
            A.ck = ck ? P.ck_state + jd.ck_off + (uint64_t)(j / K - 1) * PM : nullptr;
            A.tb_col = nullptr; A.colrec_col = colrec + (uint64_t)j * C;
            A.sn = sn; A.last = last; A.j = j; A.n = n; A.q = read[j - 1];
            A.track = j >= track_from; A.lastcol = j == n;
            column_wide<W>(sc, A, S);   // begins and ends with a CTA barrier
            if (ck) for (uint32_t a = tid; a < C; a += T) {
                CkSum cs; cs.Sm = S.Sm[a]; cs.slm = S.slm[a]; cs.tbm = S.tbm[a]; cs.pad = 0;
                P.ck_sum[jd.cksum_off + (uint64_t)(j / K - 1) * C + a] = cs;
            }
        }
        if (tid == 0) {
            int32_t g = S.cm[0];
            for (uint32_t a = 1; a < C; ++a) g = S.cm[a] > g ? S.cm[a] : g;
            gcol[n] = g;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// end-of-read fix-up, one block per read, one thread per contig-strand
// ---------------------------------------------------------------------------------------------
__global__ void fixup_kernel(const Params P) {
    const uint32_t job = P.order[blockIdx.x];
    const JobDesc jd = P.jobs[job];
    const LayoutDesc ld = P.layouts[jd.layout];
    const ContigEntry *ent = P.ents + ld.ent_off;
    const uint32_t n = jd.n;
    __shared__ int32_t s_gmax;
    __shared__ uint32_t s_first;
    // Was the y-suffix tracking window wide enough?  (dp_core.h: first_candidate_column)
    const uint32_t eff_track_from = jd.track_from > jd.j0 ? jd.track_from : jd.j0 + 1;
    const bool windowed = P.tracked_mode && !P.force_full && eff_track_from > 1 && jd.LB == 0;   // the packed tail tracks by construction
    if (windowed) {
        const int32_t *gcol = P.gcol + jd.gcol_off;
        if (threadIdx.x == 0) { s_gmax = gcol[0]; s_first = n + 1; }
        __syncthreads();
        int32_t g = gcol[0];
        for (uint32_t j = threadIdx.x; j <= n; j += blockDim.x) g = gcol[j] > g ? gcol[j] : g;
        atomicMax(&s_gmax, g);
        __syncthreads();
        const Scoring &sc = P.sc;
        int32_t submax = sc.match > sc.mismatch ? sc.match : sc.mismatch;
        if (submax < 0) submax = 0;
        const int32_t submin = sc.match < sc.mismatch ? sc.match : sc.mismatch;
        int32_t gmin = sc.g_same < sc.g_opp ? sc.g_same : sc.g_opp;
        gmin = gmin < sc.g_inter ? gmin : sc.g_inter;
        const int32_t thr = s_gmax - (submax - gmin - submin);
        uint32_t first = n + 1;
        for (uint32_t j = 1 + threadIdx.x; j <= n; j += blockDim.x) if (gcol[j] >= thr) { first = j; break; }
        atomicMin(&s_first, first);
        __syncthreads();
        if (s_first < eff_track_from) {
            if (threadIdx.x == 0) { JobOut o; o.n_chains = s_first; o.status = JOB_NEED_FULL_TRACK; P.job_out[job] = o; }
            return;
        }
    }
    if (threadIdx.x == 0) { JobOut o; o.n_chains = 0; o.status = JOB_OK; P.job_out[job] = o; }
    for (uint32_t a = threadIdx.x; a < ld.C; a += blockDim.x)
        fixup_contig(P.sc, ent[a], n, P.last + jd.cell_off, P.sn + jd.cell_off, P.tracked_mode != 0,
                     &P.colrec[jd.colrec_off + (uint64_t)n * ld.C + a].lx);
}

// ---------------------------------------------------------------------------------------------
// walk: one CTA per read.  Thread 0 walks; the CTA re-fills the unit the walk asks for.
// ---------------------------------------------------------------------------------------------
template <int W>
__device__ void refill_unit(const Params &P, const JobDesc &jd, const LayoutDesc &ld, WideSmem<W> &S, ContigEntry *s_en,
                            uint32_t a, uint32_t j, CellState *st0, CellState *st1, uint8_t *bytes, ColRec *ucr, TbUnit *unit_out) {
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t T = W * 32;
    const Scoring sc = P.sc;
    const uint32_t C = ld.C, PM = ld.PM, n = jd.n, K = P.K;
    const uint32_t b = (j - 1) / K, jb = b * K, je = (jb + K < n) ? jb + K : n;
    const ContigEntry gen = P.ents[ld.ent_off + a];
    const uint32_t pm = gen.ntiles * TILE, gbase = gen.tile_start * TILE;
    if (tid == 0) { *s_en = gen; s_en->tile_start = 0; }
    __syncthreads();
    if (b == 0) {
        column0_wide<W>(sc, s_en, nullptr, 1, gen.ntiles, st0, S, false, nullptr, n);
    } else {
        const CellState *ck = P.ck_state + jd.ck_off + (uint64_t)(b - 1) * PM + gbase;
        CellState *dst = (jb & 1u) ? st1 : st0;
        for (uint32_t p = tid; p < pm; p += T) dst[p] = ck[p];
        if (tid == 0) {
            const CkSum cs = P.ck_sum[jd.cksum_off + (uint64_t)(b - 1) * C + a];
            S.Sm[0] = cs.Sm; S.slm[0] = cs.slm; S.tbm[0] = cs.tbm;
        }
    }
    __syncthreads();
    const uint8_t *read = P.reads + jd.read_off;
    const ColRec *colrec = P.colrec + jd.colrec_off;
    for (uint32_t jj = jb + 1; jj <= je; ++jj) {
        if (tid == 0) {
            const ColRec cr = colrec[(uint64_t)jj * C + a];
            JumpInfo J; J.score = cr.jscore; J.len = cr.jlen; J.idx = cr.jidx; J.from = cr.jfrom;
            S.J[0] = J;
        }
        ColWide A;
        A.ent = s_en; A.owner = nullptr; A.C = 1; A.NT = gen.ntiles; A.bases = P.contig_bases;
        A.prev = ((jj - 1) & 1u) ? st1 : st0; A.curr = (jj & 1u) ? st1 : st0; A.ck = nullptr;
        A.tb_col = bytes + (uint64_t)(jj - jb - 1) * pm; A.colrec_col = ucr + (jj - jb - 1);
        A.sn = nullptr; A.last = nullptr; A.j = jj; A.n = n; A.q = read[jj - 1];
        A.track = false; A.lastcol = false;
        column_wide<W>(sc, A, S);
    }
    if (tid == 0) { unit_out->bytes = bytes; unit_out->cr = ucr; unit_out->a = a; unit_out->jb = jb; unit_out->je = je; unit_out->pm = pm; unit_out->i_hi = 0xffffffffu; unit_out->slope = 0; }
    __syncthreads();
}

enum : uint32_t { WCMD_DONE = 0, WCMD_UNIT = 1 };

}  // namespace gpu
}  // namespace stitch
