// kernels_walk.cuh — end-contig selection + pointer walk (traceback/mod.rs:129-373 of the reference).
//
//   walk_job             thread 0 walks; whenever the walk needs packed traceback bytes that are not loaded the whole
//                        CTA re-fills that unit (one contig over one block of K columns) from the column-state
//                        checkpoint before it, with the packed-key columns (kernels_packed.cuh) or, for scorings
//                        outside their regime, the wide ones.
//   walk_kernel          walk_job for the reads of the wide path (their fill ran in fill_wide_kernel).
//   pk_walk_phase        walk_job for the reads of the packed path, as the second phase of fill_packed_kernel.
#pragma once
#include "kernels_packed.cuh"
#include "kernels_wide.cuh"

namespace stitch {
namespace gpu {

struct WalkShared {
    uint32_t cmd, ua, uj, ui;
    ContigEntry en;
    TbUnit unit;
    PkColConst cc;
    uint8_t seen[MAX_STRANDS];
};

struct WalkBufs {           // per CTA
    CellState *st0, *st1;   // wide re-fill state
    int32_t *wps; bool wps_smem;   // packed re-fill state
    uint8_t *ubytes; ColRec *ucr;
};

// `jd` carries the offsets of the read's records.
template <int W, bool PACKED_ONLY>
__device__ __noinline__ void walk_job(const Params &P, uint32_t job, const JobDesc &jd, const LayoutDesc &ld, WideSmem<W> *WS, PackSmem &PS,
                         UnitStage &US, WalkShared &sh, const WalkBufs &B) {
    const uint32_t tid = threadIdx.x;
    const long long t_job0 = clock64();
    // thread-0 state
    ReadView v; WalkState ws; ChainHdr h;
    uint32_t used = 0, n_chains = 0, status = WALK_OK, n_seen = 0, a_cur = 0;
    bool walking = false, finished = false;
    OutOp *ops = P.ops + jd.ops_off;
    ChainHdr *hdr = P.chains + jd.chain_first;
    if (tid == 0) {
        v.sc = P.sc; v.ent = P.ents + ld.ent_off; v.C = ld.C; v.n = jd.n;
        v.colrec = P.colrec + jd.colrec_off; v.last = P.last + jd.cell_off; v.sn = P.sn + jd.cell_off;
        v.contig_bases = P.contig_bases; v.read = P.reads + jd.read_off;
        v.unit.bytes = nullptr; v.unit.cr = nullptr; v.unit.a = 0xffffffffu; v.unit.jb = v.unit.je = v.unit.pm = 0; v.unit.i_hi = 0xffffffffu; v.unit.slope = 0;
        if (jd.walk == host::WALK_ALL) for (uint32_t a = 0; a < ld.C; ++a) sh.seen[a] = 0;
    }
    for (;;) {
        if (tid == 0) {
            sh.cmd = WCMD_DONE;
            while (!finished) {
                if (!walking) {   // choose the next chain to walk
                    int a_end = -1;
                    if (jd.walk == host::WALK_BEST) { if (n_chains == 0 && used == 0) a_end = (int)pick_end(v, nullptr); }
                    else if (jd.walk == host::WALK_FROM) {
                        if (n_chains == 0 && used == 0) a_end = v.pos_of(jd.from_contig);
                    } else if (n_seen < ld.C) a_end = (int)pick_end(v, sh.seen);
                    if (a_end < 0) { finished = true; break; }
                    a_cur = (uint32_t)a_end;
                    walk_begin(v, a_cur, ops + used, jd.ops_cap - used, ws, h);
                    walking = true;
                }
                const uint32_t s = walk_run(v, ws, h);
                if (s == WALK_NEED_UNIT) { sh.cmd = WCMD_UNIT; sh.ua = ws.a; sh.uj = ws.j; sh.ui = ws.i; break; }
                walking = false;
                auto mark = [&](uint32_t idx) {
                    const int p = v.pos_of(idx);
                    if (p >= 0 && !sh.seen[p]) { sh.seen[p] = 1; ++n_seen; }
                };
                if (jd.walk == host::WALK_ALL) {
                    if (s == WALK_NONE) { mark(v.ent[a_cur].contig_idx); continue; }
                    if (s != WALK_OK) { status = s; finished = true; break; }
                    mark(h.start_contig_idx); mark(h.end_contig_idx);
                    for (uint32_t k = 0; k < h.n_ops; ++k) if (ops[used + k].kind == OP_XJUMP) mark(ops[used + k].a);
                    if (n_chains >= jd.max_chains) { status = WALK_OVERFLOW; finished = true; break; }
                    hdr[n_chains++] = h;
                    used += h.n_ops;
                } else {
                    if (s == WALK_OK) { hdr[0] = h; n_chains = 1; }
                    else if (s != WALK_NONE) status = s;
                    finished = true;
                }
            }
        }
        __syncthreads();
        if (sh.cmd == WCMD_DONE) break;
        const long long t_r0 = clock64();
        if (PACKED_ONLY || jd.LB)
            pk_refill_unit<W>(P, jd, ld, PS, US, &sh.en, &sh.cc, sh.ua, sh.uj, sh.ui, B.wps, P.wpstate_half, B.wps_smem, B.ubytes, B.ucr, &sh.unit);
        else if (!PACKED_ONLY)
            refill_unit<W>(P, jd, ld, *WS, &sh.en, sh.ua, sh.uj, B.st0, B.st1, B.ubytes, B.ucr, &sh.unit);
        if (tid == 0) {
            v.unit = sh.unit;
            if (P.dbg) {
                atomicAdd(P.dbg + 3, 1ull);
                atomicAdd(P.dbg + 4, (unsigned long long)(clock64() - t_r0));
                atomicAdd(P.dbg + 6, (unsigned long long)(sh.unit.je - sh.unit.jb));
            }
        }
    }
    if (tid == 0) {
        JobOut o; o.n_chains = n_chains; o.status = status; P.job_out[job] = o;
        if (P.dbg) atomicAdd(P.dbg + 5, (unsigned long long)(clock64() - t_job0));
    }
    __syncthreads();
}

template <int W>
__global__ void __launch_bounds__(W * 32) walk_kernel(const Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WideSmem<W> S; S.carve(smem_raw, 1);
    PackSmem PS; PS.carve(smem_raw, PackSmem::layout(1, P.max_ctiles, W, 0), 1);   // same bytes: a job uses one of the two
    UnitStage US; US.carve(smem_raw + P.walk_stage_smem_off, P.K);
    __shared__ uint32_t sJob;
    __shared__ WalkShared sh;
    const uint32_t tid = threadIdx.x;
    WalkBufs B;
    B.st0 = P.state + (uint64_t)blockIdx.x * P.state_stride; B.st1 = B.st0 + P.state_half;
    B.ubytes = P.unit_bytes + (uint64_t)blockIdx.x * P.unit_stride;
    B.ucr = P.unit_cr + (uint64_t)blockIdx.x * P.K;
    // packed state of the unit's contig: in shared memory when it fits (a contig is a few tiles), else global
    B.wps_smem = P.walk_state_smem_off != 0;
    B.wps = B.wps_smem ? reinterpret_cast<int32_t *>(smem_raw + P.walk_state_smem_off) : P.wpstate + (uint64_t)blockIdx.x * P.wpstate_stride;
    for (;;) {
        __syncthreads();
        if (tid == 0) sJob = atomicAdd(P.counter, 1u);
        __syncthreads();
        if (sJob >= P.n_jobs) break;
        const uint32_t job = P.order[sJob];
        const JobDesc jd = P.jobs[job];
        const LayoutDesc ld = P.layouts[jd.layout];
        walk_job<W, false>(P, job, jd, ld, &S, PS, US, sh, B);   // (this kernel does nothing else: no copies needed)
    }
}

// Walk phase of fill_packed_kernel (see there).  Shared memory: the tables of PackSmem as carved by the kernel, the
// per-unit staging area after them, and the re-fill state of one contig in the (now idle) stage ring.
template <int W>
__device__ __noinline__ void pk_walk_phase(const Params P, unsigned char *smem_raw) {   // by value: the caller's P stays in the constant bank
    PackSmem PS; PS.carve(smem_raw, P.pso, P.cmax);
    UnitStage US; US.carve(smem_raw + P.walk_stage_smem_off, P.K);
    __shared__ uint32_t sJob;
    __shared__ WalkShared sh;
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t T = W * 32;
    WalkBufs B;
    B.st0 = nullptr; B.st1 = nullptr;
    B.ubytes = P.unit_bytes + (uint64_t)blockIdx.x * P.unit_stride;
    B.ucr = P.unit_cr + (uint64_t)blockIdx.x * P.K;
    B.wps_smem = 2ull * P.wpstate_half * sizeof(int32_t) <= (uint64_t)P.stage_bytes;
    B.wps = B.wps_smem ? reinterpret_cast<int32_t *>(PS.stage) : P.wpstate + (uint64_t)blockIdx.x * P.wpstate_stride;
    for (;;) {
        __syncthreads();
        if (tid == 0) sJob = atomicAdd(P.counter + 1, 1u);
        __syncthreads();
        if (sJob >= P.n_jobs) break;
        const uint32_t job = P.order[sJob];
        if (tid == 0) {
            while (*reinterpret_cast<volatile uint32_t *>(P.done + job) == 0u) __nanosleep(2000);
            __threadfence();
        }
        __syncthreads();
        const JobDesc jd = P.jobs[job];
        const LayoutDesc ld = P.layouts[jd.layout];
        // end-of-read fix-up (SCA:453-555), one thread per contig-strand
        for (uint32_t a = tid; a < ld.C; a += T)
            fixup_contig(P.sc, P.ents[ld.ent_off + a], jd.n, P.last + jd.cell_off, P.sn + jd.cell_off, P.tracked_mode != 0,
                         &P.colrec[jd.colrec_off + (uint64_t)jd.n * ld.C + a].lx);
        __syncthreads();
        walk_job<W, true>(P, job, jd, ld, nullptr, PS, US, sh, B);
    }
}

}  // namespace gpu
}  // namespace stitch
