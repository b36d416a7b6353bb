// kernels_walk.cuh — end-contig selection + pointer walk (traceback/mod.rs:129-373 of the reference), one CTA
// per read: thread 0 walks; whenever the walk needs packed traceback bytes that are not loaded the whole CTA
// re-fills that unit (one contig over one block of K columns) from the column-state checkpoint before it,
// with the packed-key columns (kernels_packed.cuh) or, for scorings outside their regime, the wide ones.
#pragma once
#include "kernels_packed.cuh"
#include "kernels_wide.cuh"

namespace stitch {
namespace gpu {

template <int W>
__global__ void __launch_bounds__(W * 32) walk_kernel(const Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WideSmem<W> S; S.carve(smem_raw, 1);
    PackSmem PS; PS.carve(smem_raw, 1, P.max_ctiles, W, false);   // same bytes: a job uses one of the two
    __shared__ PkColConst s_cc;
    UnitStage US; US.carve(smem_raw + P.walk_stage_smem_off, P.K);
    __shared__ uint32_t sJob, sCmd, sUa, sUj;
    __shared__ ContigEntry s_en;
    __shared__ TbUnit s_unit;
    __shared__ uint8_t s_seen[MAX_STRANDS];
    const uint32_t tid = threadIdx.x;
    CellState *st0 = P.state + (uint64_t)blockIdx.x * P.state_stride;
    CellState *st1 = st0 + P.state_half;
    uint8_t *ubytes = P.unit_bytes + (uint64_t)blockIdx.x * P.unit_stride;
    ColRec *ucr = P.unit_cr + (uint64_t)blockIdx.x * P.K;
    // packed state of the unit's contig: in shared memory when it fits (a contig is a few tiles), else global
    int32_t *wps = P.walk_state_smem_off ? reinterpret_cast<int32_t *>(smem_raw + P.walk_state_smem_off)
                                         : P.wpstate + (uint64_t)blockIdx.x * P.wpstate_stride;

    for (;;) {
        __syncthreads();
        if (tid == 0) sJob = atomicAdd(P.counter, 1u);
        __syncthreads();
        if (sJob >= P.n_jobs) break;
        const uint32_t job = P.order[sJob];
        const JobDesc jd = P.jobs[job];
        const LayoutDesc ld = P.layouts[jd.layout];
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
            v.contig_bases = P.contig_bases; v.read = P.reads + jd.read_off; v.pos_of = P.posof + ld.posof_off;
            v.unit.bytes = nullptr; v.unit.cr = nullptr; v.unit.a = 0xffffffffu; v.unit.jb = v.unit.je = v.unit.pm = 0;
            if (jd.walk == host::WALK_ALL) for (uint32_t a = 0; a < ld.C; ++a) s_seen[a] = 0;
        }
        for (;;) {
            if (tid == 0) {
                sCmd = WCMD_DONE;
                while (!finished) {
                    if (!walking) {   // choose the next chain to walk
                        int a_end = -1;
                        if (jd.walk == host::WALK_BEST) { if (n_chains == 0 && used == 0) a_end = (int)pick_end(v, nullptr); }
                        else if (jd.walk == host::WALK_FROM) {
                            if (n_chains == 0 && used == 0) a_end = jd.from_contig < MAX_STRANDS ? v.pos_of[jd.from_contig] : -1;
                        } else if (n_seen < ld.C) a_end = (int)pick_end(v, s_seen);
                        if (a_end < 0) { finished = true; break; }
                        a_cur = (uint32_t)a_end;
                        walk_begin(v, a_cur, ops + used, jd.ops_cap - used, ws, h);
                        walking = true;
                    }
                    const uint32_t s = walk_run(v, ws, h);
                    if (s == WALK_NEED_UNIT) { sCmd = WCMD_UNIT; sUa = ws.a; sUj = ws.j; break; }
                    walking = false;
                    auto mark = [&](uint32_t idx) {
                        const int p = idx < MAX_STRANDS ? v.pos_of[idx] : -1;
                        if (p >= 0 && !s_seen[p]) { s_seen[p] = 1; ++n_seen; }
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
            if (sCmd == WCMD_DONE) break;
            const long long t_r0 = clock64();
            if (jd.LB) pk_refill_unit<W>(P, jd, ld, PS, US, &s_en, &s_cc, sUa, sUj, wps, P.wpstate_half, P.walk_state_smem_off != 0, ubytes, ucr, &s_unit);
            else refill_unit<W>(P, jd, ld, S, &s_en, sUa, sUj, st0, st1, ubytes, ucr, &s_unit);
            if (tid == 0) {
                v.unit = s_unit;
                if (P.dbg) {
                    atomicAdd(P.dbg + 3, 1ull);
                    atomicAdd(P.dbg + 4, (unsigned long long)(clock64() - t_r0));
                    atomicAdd(P.dbg + 6, (unsigned long long)(s_unit.je - s_unit.jb));
                }
            }
        }
        if (tid == 0) {
            JobOut o; o.n_chains = n_chains; o.status = status; P.job_out[job] = o;
            if (P.dbg) atomicAdd(P.dbg + 5, (unsigned long long)(clock64() - t_job0));
        }
    }
}

}  // namespace gpu
}  // namespace stitch
