// cuda_backend.cu — the sm_100a implementation of the DP backend + the product's C ABI.
//
// Kernels (one launch each per chunk of reads):
//   fill_kernel    K1  column-by-column DP fill of every contig-strand of a read; persistent CTAs pull
//                      reads from a queue; replaces MCA::custom's loop body (multi_contig_aligner.rs:270-347
//                      = SCA:188-239 + 677-697 + 292-451 of the reference)
//   fixup_kernel   K2  end-of-read fix-up of column n (SCA:453-555), one thread per contig-strand
//   walk_kernel    K3  end-contig selection + pointer walk (traceback/mod.rs:129-373)
// The arithmetic lives in dp_core.h (shared with the CPU emulator the tests fuzz against the
// oracle); this file is the parallel glue: tile/lane mapping, warp-shuffle max-plus scan of the
// insertion chain, cross-warp carry folding through shared memory, per-contig reductions.
//
// There is no CPU path: every entry point fails with STITCH_ERR_CUDA when no device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <numeric>

#define STITCH_API(name) stitch_##name
#include "capi_impl.hpp"

namespace stitch {
namespace gpu {

using host::Error;

#define CUDA_CHECK(expr)                                                                              \
    do {                                                                                              \
        cudaError_t err__ = (expr);                                                                   \
        if (err__ != cudaSuccess)                                                                     \
            throw Error(err__ == cudaErrorMemoryAllocation ? STITCH_ERR_NOMEM : STITCH_ERR_CUDA,      \
                        std::string(#expr) + ": " + cudaGetErrorString(err__));                       \
    } while (0)

constexpr int FILL_WARPS = 8;
constexpr unsigned FULL = 0xffffffffu;

struct JobDesc {
    uint64_t read_off;     // into the reads blob
    uint64_t tb_off;       // bytes into the traceback arena
    uint64_t colrec_off;   // ColRec records
    uint64_t cell_off;     // LastCell / SnRec records
    uint64_t ops_off;      // OutOp records
    uint32_t n, layout;
    uint32_t walk, from_contig;
    uint32_t ops_cap, chain_first, max_chains, pad;
};

struct LayoutDesc { uint32_t ent_off, C, n_tiles, owner_off, posof_off, PM, pad0, pad1; };

struct JobOut { uint32_t n_chains, status; };

struct Params {
    Scoring sc;
    const JobDesc *jobs;
    const uint32_t *order;
    uint32_t n_jobs, cmax;
    const LayoutDesc *layouts;
    const ContigEntry *ents;
    const uint16_t *owners;
    const int16_t *posof;
    const uint8_t *contig_bases;
    const uint8_t *reads;
    CellState *state;
    uint64_t state_stride;   // CellState records per CTA (two column buffers)
    uint64_t state_half;     // records per column buffer
    uint8_t *tb;
    ColRec *colrec;
    LastCell *last;
    SnRec *sn;
    OutOp *ops;
    ChainHdr *chains;
    JobOut *job_out;
    uint32_t *counter;
    int track;
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

static size_t fill_smem_bytes(uint32_t cmax, int W) {
    size_t b = 0;
    b += sizeof(JumpInfo) * cmax;
    b += sizeof(RowM) * cmax;
    b += sizeof(XsPart) * cmax * W;
    b += sizeof(CmPart) * cmax * W;
    b += sizeof(int32_t) * cmax * 6;
    return b + 64;
}

// ---------------------------------------------------------------------------------------------
// K1: fill
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(W * 32) fill_kernel(const Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t cmax = P.cmax;
    JumpInfo *sJ = reinterpret_cast<JumpInfo *>(smem_raw);
    RowM *sRowM = reinterpret_cast<RowM *>(sJ + cmax);
    XsPart *sXs = reinterpret_cast<XsPart *>(sRowM + cmax);
    CmPart *sCmp = reinterpret_cast<CmPart *>(sXs + (size_t)cmax * W);
    int32_t *sCm = reinterpret_cast<int32_t *>(sCmp + (size_t)cmax * W);
    uint32_t *sCml = reinterpret_cast<uint32_t *>(sCm + cmax);
    uint32_t *sCmk = sCml + cmax;
    int32_t *sSm = reinterpret_cast<int32_t *>(sCmk + cmax);
    uint32_t *sSlm = reinterpret_cast<uint32_t *>(sSm + cmax);
    uint32_t *sTbm = sSlm + cmax;
    __shared__ ICarry sTileAgg[2][W];
    __shared__ ICarry sRound[2];
    __shared__ uint32_t sJob;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr uint32_t T = W * 32;
    const Scoring sc = P.sc;
    const bool track = P.track != 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) sJob = atomicAdd(P.counter, 1u);
        __syncthreads();
        if (sJob >= P.n_jobs) break;
        const JobDesc jd = P.jobs[P.order[sJob]];
        const LayoutDesc ld = P.layouts[jd.layout];
        const ContigEntry *ent = P.ents + ld.ent_off;
        const uint16_t *owner = P.owners + ld.owner_off;
        const uint32_t C = ld.C, NT = ld.n_tiles, PM = ld.PM, n = jd.n;
        CellState *st0 = P.state + (uint64_t)blockIdx.x * P.state_stride;
        CellState *st1 = st0 + P.state_half;
        uint8_t *tb = P.tb + jd.tb_off;
        ColRec *colrec = P.colrec + jd.colrec_off;
        LastCell *last = P.last + jd.cell_off;
        SnRec *sn = P.sn + jd.cell_off;
        const uint8_t *read = P.reads + jd.read_off;

        // ---- column 0 (SCA:97-186) ----
        for (uint32_t tile = warp; tile < NT; tile += W) {
            const uint32_t a = owner[tile];
            const ContigEntry en = ent[a];
            const uint32_t tic = tile - en.tile_start;
            STITCH_UNROLL
            for (int k = 0; k < STRIP; ++k) {
                const uint32_t i = tic * TILE + lane * STRIP + (uint32_t)k + 1;
                if (i <= en.m) {
                    const Col0 c0 = col0_at(sc, i, en.m);
                    CellState s; s.S = c0.S; s.D = MIN_SCORE; s.sl = c0.sl; s.dl = 0;
                    const uint32_t si = state_index(tile, lane, (uint32_t)k);
                    st0[si] = s;
                    if (track) sn[si] = sn_init(sc, c0.S, c0.sl, en.contig_idx, n);
                }
            }
        }
        for (uint32_t a = tid; a < C; a += T) {
            const ContigEntry en = ent[a];
            sCm[a] = 0; sCml[a] = 0; sCmk[a] = 0;   // column-0 best is S(0,0) = 0 at row 0
            const Col0 cm = col0_at(sc, en.m, en.m);
            sSm[a] = cm.S; sSlm[a] = cm.sl; sTbm[a] = cm.s_tb;
            int32_t t; uint32_t lx; col0_tracker(sc, en.m, t, lx);
            ColRec cr; cr.jidx = 0; cr.jfrom = 0; cr.lx = lx; cr.pad = 0;
            colrec[a] = cr;
        }
        __syncthreads();

        for (uint32_t j = 1; j <= n; ++j) {
            const CellState *prev = (j & 1u) ? st0 : st1;
            CellState *curr = (j & 1u) ? st1 : st0;
            const Row0 r0 = row0_at(sc, j, n), r0p = row0_at(sc, j - 1, n);
            ColConst cc; cc.j = j; cc.n = n; cc.q = read[j - 1];
            { const int32_t dj = sc.o + sc.e * (int32_t)j; cc.xclip_score = sc.xp + (sc.yp > dj ? sc.yp : dj); }
            cc.sl0j = r0.sl;
            uint8_t *tb_col = tb + (uint64_t)(j - 1) * PM;
            const bool lastcol = (j == n);

            // ---- jump selection for this column (MCA:279-331) ----
            for (uint32_t a = tid; a < C; a += T) sJ[a] = select_jump(sc, ent, C, a, sCm, sCml, sCmk);
            for (uint32_t x = tid; x < C * W; x += T) { xs_init(sXs[x]); cm_init(sCmp[x]); }
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
                    a = owner[tile];
                    const ContigEntry en = ent[a];
                    tc.a = a; tc.self_idx = en.contig_idx; tc.m = en.m; tc.tile = tile; tc.tile_in_contig = tile - en.tile_start;
                    tc.J = sJ[a]; tc.circular = en.circular != 0; tc.wrap_src_ok = sTbm[a] != TB_XCLIP_SUFFIX;
                    tc.Sm_prev = sSm[a]; tc.slm_prev = sSlm[a];
                    row0 = tc.tile_in_contig * TILE + lane * STRIP + 1;
                    lact = row0 <= en.m;
                    CellState up[STRIP];
                    STITCH_UNROLL
                    for (int k = 0; k < STRIP; ++k) {
                        if (lact) {
                            const int4 v = *reinterpret_cast<const int4 *>(prev + state_index(tile, lane, (uint32_t)k));
                            up[k].S = v.x; up[k].D = v.y; up[k].sl = (uint32_t)v.z; up[k].dl = (uint32_t)v.w;
                        } else { up[k].S = MIN_SCORE; up[k].D = MIN_SCORE; up[k].sl = 0; up[k].dl = 0; }
                    }
                    if (lact) {
                        const uint8_t *xb = P.contig_bases + en.seq_off + row0 - 1;
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
                            const CellState c = prev[state_index(tile - 1, 31, STRIP - 1)];
                            dgS = c.S; dgsl = c.sl;
                        }
                    }
                    if (lact) lane_pass_a(sc, cc, tc, row0, up, dgS, dgsl, x, la, &sRowM[a]);
                    // inclusive max-plus scan of the lane aggregates
                    ICarry inc = la.agg;
                    STITCH_UNROLL
                    for (int d = 1; d < 32; d <<= 1) {
                        const ICarry o = shfl_up_ic(inc, d);
                        if ((int)lane >= d) inc = icarry_combine(o, (uint32_t)(d * STRIP), sc.e, inc);
                    }
                    excl = shfl_up_ic(inc, 1);
                    if (lane == 31) sTileAgg[par][warp] = inc;
                }
                __syncthreads();
                if (tact) {
                    // carry into this tile: fold the aggregates of the tiles of the same contig before it
                    uint32_t w0 = warp;
                    while (w0 > 0 && owner[t0 + w0 - 1] == a) --w0;
                    ICarry c;
                    if (t0 + w0 == ent[a].tile_start) c = icarry_row1(sc, r0);
                    else c = sRound[par];
                    for (uint32_t u = w0; u < warp; ++u) c = icarry_combine(c, TILE, sc.e, sTileAgg[par][u]);
                    const uint32_t last_tile = (t0 + W < NT ? t0 + W : NT) - 1;
                    if (tile == last_tile && lane == 0) sRound[par ^ 1u] = icarry_combine(c, TILE, sc.e, sTileAgg[par][warp]);
                    const ICarry cin = lane == 0 ? c : icarry_combine(c, lane * STRIP, sc.e, excl);
                    LaneB lb; xs_init(lb.xs); cm_init(lb.cm);
                    if (lact) lane_pass_b(sc, cc, tc, row0, lane, la, cin, curr, tb_col, track, sn, lastcol, last, x, lb, &sRowM[a]);
                    STITCH_UNROLL
                    for (int d = 16; d >= 1; d >>= 1) {
                        lb.xs = xs_merge(lb.xs, shfl_xor_xs(lb.xs, d));
                        lb.cm = cm_merge(lb.cm, shfl_xor_cm(lb.cm, d));
                    }
                    if (lane == 0) {
                        const uint32_t slot = a * W + warp;
                        sXs[slot] = xs_merge(sXs[slot], lb.xs);
                        sCmp[slot] = cm_merge(sCmp[slot], lb.cm);
                    }
                }
            }
            __syncthreads();

            // ---- per contig: finish row m, column best for the next jump ----
            for (uint32_t a = warp; a < C; a += W) {
                XsPart xs; CmPart cm; xs_init(xs); cm_init(cm);
                if (lane < (uint32_t)W) { xs = sXs[a * W + lane]; cm = sCmp[a * W + lane]; }
                STITCH_UNROLL
                for (int d = 16; d >= 1; d >>= 1) {
                    xs = xs_merge(xs, shfl_xor_xs(xs, d));
                    cm = cm_merge(cm, shfl_xor_cm(cm, d));
                }
                if (lane == 0) {
                    const ContigColOut o = contig_finalize(sc, cc, ent[a], a, C, sRowM[a], xs, cm, r0, sJ[a], curr, tb_col,
                                                           colrec + (uint64_t)j * C, track, sn, lastcol, last);
                    sCm[a] = o.cm.S; sCmk[a] = o.cm.row; sCml[a] = o.cm.sl;
                    sSm[a] = o.Sm; sSlm[a] = o.slm; sTbm[a] = o.s_tb_m;
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2: end-of-read fix-up, one thread per (read, contig-strand)
// ---------------------------------------------------------------------------------------------
__global__ void fixup_kernel(const Params P) {
    const uint32_t job = blockIdx.x;
    const JobDesc jd = P.jobs[job];
    const LayoutDesc ld = P.layouts[jd.layout];
    const ContigEntry *ent = P.ents + ld.ent_off;
    for (uint32_t a = threadIdx.x; a < ld.C; a += blockDim.x)
        fixup_contig(P.sc, ent[a], jd.n, P.last + jd.cell_off, P.sn + jd.cell_off, P.track != 0,
                     &P.colrec[jd.colrec_off + (uint64_t)jd.n * ld.C + a].lx);
}

// ---------------------------------------------------------------------------------------------
// K3: traceback walks, one thread per read
// ---------------------------------------------------------------------------------------------
__global__ void walk_kernel(const Params P) {
    const uint32_t job = blockIdx.x * blockDim.x + threadIdx.x;
    if (job >= P.n_jobs) return;
    const JobDesc jd = P.jobs[job];
    const LayoutDesc ld = P.layouts[jd.layout];
    ReadView v;
    v.sc = P.sc; v.ent = P.ents + ld.ent_off; v.C = ld.C; v.n = jd.n; v.PM = ld.PM;
    v.tb = P.tb + jd.tb_off; v.colrec = P.colrec + jd.colrec_off; v.last = P.last + jd.cell_off; v.sn = P.sn + jd.cell_off;
    v.contig_bases = P.contig_bases; v.read = P.reads + jd.read_off; v.pos_of = P.posof + ld.posof_off;
    OutOp *ops = P.ops + jd.ops_off;
    ChainHdr *hdr = P.chains + jd.chain_first;
    uint32_t used = 0, n_chains = 0, status = WALK_OK;
    if (jd.walk == host::WALK_BEST || jd.walk == host::WALK_FROM) {
        int a_end = -1;
        if (jd.walk == host::WALK_BEST) a_end = (int)pick_end(v, nullptr);
        else a_end = jd.from_contig < MAX_STRANDS ? v.pos_of[jd.from_contig] : -1;
        if (a_end >= 0) {
            ChainHdr h;
            walk_chain(v, (uint32_t)a_end, ops, jd.ops_cap, h);
            if (h.status == WALK_OK) { hdr[0] = h; n_chains = 1; }
            else if (h.status != WALK_NONE) status = h.status;
        }
    } else {
        uint8_t seen[MAX_STRANDS];
        for (uint32_t a = 0; a < ld.C; ++a) seen[a] = 0;
        uint32_t n_seen = 0;
        while (n_seen < ld.C && status == WALK_OK) {
            const uint32_t a_end = pick_end(v, seen);
            ChainHdr h;
            walk_chain(v, a_end, ops + used, jd.ops_cap - used, h);
            auto mark = [&](uint32_t idx) {
                const int p = idx < MAX_STRANDS ? v.pos_of[idx] : -1;
                if (p >= 0 && !seen[p]) { seen[p] = 1; ++n_seen; }
            };
            if (h.status == WALK_NONE) { mark(v.ent[a_end].contig_idx); continue; }
            if (h.status != WALK_OK) { status = h.status; break; }
            mark(h.start_contig_idx); mark(h.end_contig_idx);
            for (uint32_t k = 0; k < h.n_ops; ++k) if (ops[used + k].kind == OP_XJUMP) mark(ops[used + k].a);
            if (n_chains >= jd.max_chains) { status = WALK_OVERFLOW; break; }
            hdr[n_chains++] = h;
            used += h.n_ops;
        }
    }
    JobOut o; o.n_chains = n_chains; o.status = status;
    P.job_out[job] = o;
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
    T *p = nullptr; size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = n + n / 8 + 64;
        CUDA_CHECK(cudaMalloc(&p, want * sizeof(T)));
        cap = want;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    ~DevBuf() { release(); }
};
template <typename T>
struct PinBuf {
    T *p = nullptr; size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = n + n / 8 + 64;
        CUDA_CHECK(cudaMallocHost(&p, want * sizeof(T)));
        cap = want;
    }
    ~PinBuf() { if (p) cudaFreeHost(p); }
};

static inline uint64_t round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

struct CudaBackend : host::Backend {
    host::Aligner &al;
    int device;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int num_sms = 0;
    uint32_t max_inflight = 0;
    size_t uploaded_layouts = 0;
    uint32_t cmax = 1;
    const uint8_t *device_reads = nullptr;   // set for run_device()

    DevBuf<uint8_t> d_contigs, d_reads, d_tb;
    DevBuf<ContigEntry> d_ents;
    DevBuf<uint16_t> d_owners;
    DevBuf<int16_t> d_posof;
    DevBuf<LayoutDesc> d_layouts;
    DevBuf<JobDesc> d_jobs;
    DevBuf<uint32_t> d_order;
    DevBuf<CellState> d_state;
    DevBuf<ColRec> d_colrec;
    DevBuf<LastCell> d_last;
    DevBuf<SnRec> d_sn;
    DevBuf<OutOp> d_ops;
    DevBuf<ChainHdr> d_chains;
    DevBuf<JobOut> d_jobout;
    DevBuf<uint32_t> d_counter;
    PinBuf<uint8_t> h_reads;
    PinBuf<JobDesc> h_jobs;
    PinBuf<uint32_t> h_order;
    PinBuf<OutOp> h_ops;
    PinBuf<ChainHdr> h_chains;
    PinBuf<JobOut> h_jobout;

    CudaBackend(host::Aligner &a, int dev) : al(a), device(dev) {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            throw Error(STITCH_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                             "); stitch_b200 has no CPU fallback");
        if (dev < 0 || dev >= count) throw Error(STITCH_ERR_INVALID, "bad device index");
        CUDA_CHECK(cudaSetDevice(dev));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
        num_sms = prop.multiProcessorCount;
        CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (auto &x : ev) CUDA_CHECK(cudaEventCreate(&x));
        d_contigs.reserve(al.contigs.blob.size() + 64);
        CUDA_CHECK(cudaMemcpy(d_contigs.p, al.contigs.blob.data(), al.contigs.blob.size(), cudaMemcpyHostToDevice));
        d_counter.reserve(1);
    }
    ~CudaBackend() override {
        cudaSetDevice(device);
        for (auto &x : ev) if (x) cudaEventDestroy(x);
        if (stream) cudaStreamDestroy(stream);
    }
    void set_max_inflight(uint32_t n) override { max_inflight = n; }

    void upload_layouts() {
        const auto &Ls = al.layouts.layouts;
        if (uploaded_layouts == Ls.size()) return;
        std::vector<ContigEntry> ents; std::vector<uint16_t> owners; std::vector<int16_t> posof; std::vector<LayoutDesc> descs;
        cmax = 1;
        for (const auto &L : Ls) {
            LayoutDesc d{};
            d.ent_off = (uint32_t)ents.size(); d.C = (uint32_t)L.ent.size(); d.n_tiles = L.n_tiles;
            d.owner_off = (uint32_t)owners.size(); d.posof_off = (uint32_t)posof.size(); d.PM = L.PM();
            cmax = std::max(cmax, d.C);
            for (uint32_t a = 0; a < d.C; ++a) {
                ents.push_back(L.ent[a]);
                for (uint32_t t = 0; t < L.ent[a].ntiles; ++t) owners.push_back((uint16_t)a);
            }
            posof.insert(posof.end(), L.pos_of.begin(), L.pos_of.end());
            descs.push_back(d);
        }
        d_ents.reserve(ents.size()); d_owners.reserve(owners.size()); d_posof.reserve(posof.size()); d_layouts.reserve(descs.size());
        CUDA_CHECK(cudaMemcpy(d_ents.p, ents.data(), ents.size() * sizeof(ContigEntry), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(d_owners.p, owners.data(), owners.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(d_posof.p, posof.data(), posof.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(d_layouts.p, descs.data(), descs.size() * sizeof(LayoutDesc), cudaMemcpyHostToDevice));
        uploaded_layouts = Ls.size();
    }

    uint64_t tb_bytes_of(const host::Job &j) const {
        return round_up((uint64_t)std::max<uint32_t>(j.n, 1) * al.layouts.layouts[j.layout].PM(), 256);
    }

    void run(const std::vector<host::Job> &jobs, std::vector<host::JobResult> &out) override {
        CUDA_CHECK(cudaSetDevice(device));
        out.assign(jobs.size(), host::JobResult());
        if (jobs.empty()) return;
        for (const auto &j : jobs) if (j.n == 0) throw Error(STITCH_ERR_INVALID, "empty read");
        upload_layouts();
        size_t free_b = 0, total_b = 0;
        CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
        // memory the traceback arena may take: what is free now plus what our own arena already holds
        const uint64_t avail = (uint64_t)free_b + d_tb.cap;
        const uint64_t budget = (uint64_t)((double)avail * 0.70);
        size_t begin = 0;
        while (begin < jobs.size()) {
            size_t end = begin; uint64_t tb = 0;
            while (end < jobs.size()) {
                const uint64_t b = tb_bytes_of(jobs[end]);
                if (end > begin && (tb + b > budget || (max_inflight && end - begin >= max_inflight))) break;
                if (b > budget) throw Error(STITCH_ERR_NOMEM, "one read's dense traceback does not fit in device memory");
                tb += b; ++end;
            }
            run_chunk(jobs, begin, end, out);
            begin = end;
        }
    }

    void run_chunk(const std::vector<host::Job> &jobs, size_t begin, size_t end, std::vector<host::JobResult> &out) {
        const uint32_t nj = (uint32_t)(end - begin);
        const auto &Ls = al.layouts.layouts;
        h_jobs.reserve(nj); h_order.reserve(nj);
        uint64_t reads_b = 0, tb_b = 0, colrec_n = 0, cell_n = 0, ops_n = 0, chains_n = 0, pm_max = 0, cells = 0;
        for (uint32_t k = 0; k < nj; ++k) {
            const host::Job &j = jobs[begin + k];
            const host::Layout &L = Ls[j.layout];
            const uint32_t C = (uint32_t)L.ent.size();
            JobDesc d{};
            d.read_off = device_reads ? (uint64_t)(uintptr_t)j.read : reads_b;
            d.tb_off = tb_b; d.colrec_off = colrec_n; d.cell_off = cell_n; d.ops_off = ops_n;
            d.n = j.n; d.layout = j.layout; d.walk = j.walk; d.from_contig = j.from_contig;
            d.max_chains = j.walk == host::WALK_ALL ? C : 1;
            const uint32_t per_chain = j.n / 2 + 4 * C + 64;
            d.ops_cap = per_chain * (j.walk == host::WALK_ALL ? std::min<uint32_t>(C, 8) : 1) * ops_scale;
            d.chain_first = (uint32_t)chains_n;
            h_jobs.p[k] = d;
            reads_b += round_up(j.n, 16); tb_b += tb_bytes_of(j);
            colrec_n += (uint64_t)(j.n + 1) * C; cell_n += L.PM(); ops_n += d.ops_cap; chains_n += d.max_chains;
            pm_max = std::max<uint64_t>(pm_max, L.PM());
            cells += L.cells_per_col * j.n;
        }
        std::iota(h_order.p, h_order.p + nj, 0u);
        std::stable_sort(h_order.p, h_order.p + nj, [&](uint32_t x, uint32_t y) {
            const host::Job &a = jobs[begin + x], &b = jobs[begin + y];
            return Ls[a.layout].cells_per_col * a.n > Ls[b.layout].cells_per_col * b.n;
        });
        const bool track = al.opts.sc.ys != MIN_SCORE;
        int ctas_per_sm = 2;
        const uint32_t grid = std::min<uint32_t>(nj, (uint32_t)(num_sms * ctas_per_sm));
        d_jobs.reserve(nj); d_order.reserve(nj); d_tb.reserve(tb_b); d_colrec.reserve(colrec_n); d_last.reserve(cell_n);
        d_sn.reserve(cell_n); d_ops.reserve(ops_n); d_chains.reserve(chains_n); d_jobout.reserve(nj);
        d_state.reserve((uint64_t)grid * 2 * pm_max);
        if (!device_reads) { d_reads.reserve(reads_b); h_reads.reserve(reads_b); }
        h_ops.reserve(ops_n); h_chains.reserve(chains_n); h_jobout.reserve(nj);

        CUDA_CHECK(cudaEventRecord(ev[0], stream));
        if (!device_reads) {
            for (uint32_t k = 0; k < nj; ++k)
                std::memcpy(h_reads.p + h_jobs.p[k].read_off, jobs[begin + k].read, jobs[begin + k].n);
            CUDA_CHECK(cudaMemcpyAsync(d_reads.p, h_reads.p, reads_b, cudaMemcpyHostToDevice, stream));
            stats.h2d += reads_b;
        }
        CUDA_CHECK(cudaMemcpyAsync(d_jobs.p, h_jobs.p, nj * sizeof(JobDesc), cudaMemcpyHostToDevice, stream));
        CUDA_CHECK(cudaMemcpyAsync(d_order.p, h_order.p, nj * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
        stats.h2d += nj * (sizeof(JobDesc) + sizeof(uint32_t));
        CUDA_CHECK(cudaMemsetAsync(d_counter.p, 0, sizeof(uint32_t), stream));

        Params P{};
        P.sc = al.opts.sc; P.jobs = d_jobs.p; P.order = d_order.p; P.n_jobs = nj; P.cmax = cmax;
        P.layouts = d_layouts.p; P.ents = d_ents.p; P.owners = d_owners.p; P.posof = d_posof.p;
        P.contig_bases = d_contigs.p; P.reads = device_reads ? device_reads : d_reads.p;
        P.state = d_state.p; P.state_stride = 2 * pm_max; P.state_half = pm_max;
        P.tb = d_tb.p; P.colrec = d_colrec.p; P.last = d_last.p; P.sn = d_sn.p; P.ops = d_ops.p; P.chains = d_chains.p;
        P.job_out = d_jobout.p; P.counter = d_counter.p; P.track = track ? 1 : 0;

        const size_t smem = fill_smem_bytes(cmax, FILL_WARPS);
        if (smem > 48 * 1024)
            CUDA_CHECK(cudaFuncSetAttribute(fill_kernel<FILL_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_CHECK(cudaEventRecord(ev[1], stream));
        fill_kernel<FILL_WARPS><<<grid, FILL_WARPS * 32, smem, stream>>>(P);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaEventRecord(ev[2], stream));
        fixup_kernel<<<nj, 64, 0, stream>>>(P);
        CUDA_CHECK(cudaGetLastError());
        walk_kernel<<<(nj + 31) / 32, 32, 0, stream>>>(P);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaEventRecord(ev[3], stream));
        CUDA_CHECK(cudaMemcpyAsync(h_jobout.p, d_jobout.p, nj * sizeof(JobOut), cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaMemcpyAsync(h_chains.p, d_chains.p, chains_n * sizeof(ChainHdr), cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaMemcpyAsync(h_ops.p, d_ops.p, ops_n * sizeof(OutOp), cudaMemcpyDeviceToHost, stream));
        cudaEvent_t ev_end;
        CUDA_CHECK(cudaEventCreate(&ev_end));
        CUDA_CHECK(cudaEventRecord(ev_end, stream));
        cudaError_t se = cudaStreamSynchronize(stream);
        if (se != cudaSuccess) { cudaEventDestroy(ev_end); throw Error(STITCH_ERR_CUDA, std::string("kernel failed: ") + cudaGetErrorString(se)); }
        float ms_fill = 0, ms_tb = 0, ms_total = 0;
        cudaEventElapsedTime(&ms_fill, ev[1], ev[2]);
        cudaEventElapsedTime(&ms_tb, ev[2], ev[3]);
        cudaEventElapsedTime(&ms_total, ev[0], ev_end);
        cudaEventDestroy(ev_end);
        stats.fill_ms += ms_fill; stats.tb_ms += ms_tb; stats.total_ms += ms_total;
        stats.launches += 3; stats.cells += cells; stats.fills += nj; stats.tb_bytes += tb_b;
        stats.d2h += nj * sizeof(JobOut) + chains_n * sizeof(ChainHdr) + ops_n * sizeof(OutOp);

        if (const char *dump = std::getenv("STITCH_DUMP_DIR")) {   // debugging aid: raw column-n / tracker records
            for (uint32_t k = 0; k < nj; ++k) {
                const JobDesc &d = h_jobs.p[k];
                const uint32_t PMk = Ls[jobs[begin + k].layout].PM(), Ck = (uint32_t)Ls[jobs[begin + k].layout].ent.size();
                std::vector<LastCell> hl(PMk); std::vector<SnRec> hs(PMk); std::vector<ColRec> hc((size_t)(d.n + 1) * Ck);
                std::vector<uint8_t> ht((size_t)d.n * PMk);
                cudaMemcpy(hl.data(), d_last.p + d.cell_off, PMk * sizeof(LastCell), cudaMemcpyDeviceToHost);
                cudaMemcpy(hs.data(), d_sn.p + d.cell_off, PMk * sizeof(SnRec), cudaMemcpyDeviceToHost);
                cudaMemcpy(hc.data(), d_colrec.p + d.colrec_off, hc.size() * sizeof(ColRec), cudaMemcpyDeviceToHost);
                cudaMemcpy(ht.data(), d_tb.p + d.tb_off, ht.size(), cudaMemcpyDeviceToHost);
                host::dump_job(dump, dump_seq++, hl, hs, hc, ht);
            }
        }
        bool overflow = false;
        for (uint32_t k = 0; k < nj; ++k) {
            if (h_jobout.p[k].status == WALK_OVERFLOW) overflow = true;
            else if (h_jobout.p[k].status == WALK_PANIC)
                throw Error(STITCH_ERR_INTERNAL, "traceback reached a state the reference panics on");
        }
        if (overflow) {
            if (ops_scale >= 64) throw Error(STITCH_ERR_INTERNAL, "operation buffer overflow");
            ops_scale *= 4;   // rare: rerun the chunk with larger operation buffers
            for (size_t k = begin; k < end; ++k) out[k] = host::JobResult();
            run_chunk(jobs, begin, end, out);
            ops_scale = 1;
            return;
        }
        for (uint32_t k = 0; k < nj; ++k) {
            const JobDesc &d = h_jobs.p[k];
            uint64_t off = d.ops_off;
            for (uint32_t c = 0; c < h_jobout.p[k].n_chains; ++c) {
                host::RawChain rc;
                rc.h = h_chains.p[d.chain_first + c];
                rc.ops.assign(h_ops.p + off, h_ops.p + off + rc.h.n_ops);
                off += rc.h.n_ops;
                out[begin + k].chains.push_back(std::move(rc));
            }
        }
    }
    uint32_t ops_scale = 1;
    uint32_t dump_seq = 0;
};

}  // namespace gpu

namespace host {
Backend *stitch_make_backend(Aligner &al, int device) { return new gpu::CudaBackend(al, device); }
}  // namespace host
}  // namespace stitch

// Reads already resident in device memory (upper-cased): times the path without host copies.
extern "C" int stitch_custom_batch_device(stitch_ctx *ctx, const uint8_t *d_bases, const uint64_t *h_offsets,
                                          uint32_t n_reads, stitch_results **out) {
    if (!ctx) return STITCH_ERR_INVALID;
    STITCH_GUARD_BEGIN
    if (!d_bases || !h_offsets || !out) throw stitch::host::Error(STITCH_ERR_INVALID, "null argument");
    auto *be = static_cast<stitch::gpu::CudaBackend *>(ctx->al.backend.get());
    std::vector<stitch::host::Job> jobs(n_reads);
    uint64_t max_n = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        jobs[r].read = reinterpret_cast<const uint8_t *>((uintptr_t)h_offsets[r]);
        jobs[r].n = (uint32_t)(h_offsets[r + 1] - h_offsets[r]);
        jobs[r].layout = ctx->al.layouts.all();
        jobs[r].walk = stitch::host::WALK_BEST;
        max_n = std::max<uint64_t>(max_n, jobs[r].n);
    }
    ctx->al.check_ranges(max_n);
    be->stats.reset();
    std::vector<stitch::host::JobResult> res;
    be->device_reads = d_bases;
    try { be->run(jobs, res); } catch (...) { be->device_reads = nullptr; throw; }
    be->device_reads = nullptr;
    std::unique_ptr<stitch_results> r(new stitch_results());
    for (uint32_t k = 0; k < n_reads; ++k) {
        r->res.begin_read();
        if (res[k].chains.size() != 1) throw stitch::host::Error(STITCH_ERR_INTERNAL, "traceback_from returned None");
        r->res.add(stitch::host::from_raw(res[k].chains[0]));
    }
    *out = r.release();
    return STITCH_OK;
    STITCH_GUARD_END(ctx)
}

// ---------------------------------------------------------------------------------------------
// INT32 add+max issue-rate microbenchmark (the roofline denominator of the fill kernel)
// ---------------------------------------------------------------------------------------------
namespace stitch { namespace gpu {
__global__ void __launch_bounds__(256) int32_peak_kernel(int32_t *out, int iters, int32_t c, int32_t w) {
    int32_t v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = (int32_t)(threadIdx.x * 16 + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = max(v[k] + c, w + k);   // one add + one max per element
        }
    }
    int32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) acc ^= v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
}}  // namespace stitch::gpu

extern "C" int stitch_measure_int32_peak(int device, double *gops) {
    if (!gops) return STITCH_ERR_INVALID;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return STITCH_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return STITCH_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return STITCH_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2048;
    int32_t *d = nullptr;
    if (cudaMalloc(&d, (size_t)blocks * threads * sizeof(int32_t)) != cudaSuccess) return STITCH_ERR_NOMEM;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        stitch::gpu::int32_peak_kernel<<<blocks, threads>>>(d, iters, -3, -1000000 + rep);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) { cudaFree(d); return STITCH_ERR_CUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double ops = 2.0 * 16 * 4 * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, ops / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    *gops = best;
    return STITCH_OK;
}
